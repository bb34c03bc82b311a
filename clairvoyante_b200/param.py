"""Hyper-parameters and tensor geometry (mirrors reference clairvoyante/param.py:1-35;
the values are the reference's interface contract, mutated at run time by the drivers)."""
NUM_THREADS = 12
maxEpoch = 10000
parameterOutputPlaceHolder = 6

# Tensor related parameters
flankingBaseNum = 16
matrixNum = 4
bloscBlockSize = 500
expandReferenceRegion = 1000000   # dataPrepScripts/param.py:3 (reference bases fetched around a --ctgStart/--ctgEnd region)

# Model hyperparameters
trainBatchSize = 10000
predictBatchSize = 1000
initialLearningRate = 0.001
learningRateDecay = 0.1
maxLearningRateSwitch = 3
trainingDatasetPercentage = 0.9

# Clairvoyante v3 specific
l2RegularizationLambda = 0.001
l2RegularizationLambdaDecay = 0.1
dropoutRateFC4 = 0.5
dropoutRateFC5 = 0.0


def str2bool(v):
    """param.py:28-35"""
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    elif v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    else:
        import sys
        raise sys.exit('Boolean value expected.')
