"""Run-time knobs shared by the drivers -- the reference's `param` module is part of its interface: callVar / train /
tensor2Bin read AND mutate these attributes (`param.NUM_THREADS = ...`, callVar.py:38-41), so names and default values are
those of clairvoyante/param.py:1-25 and dataPrepScripts/param.py:1-3; nothing else is taken from there."""

# ---- tensor geometry: (2 * flankingBaseNum + 1) positions x 4 bases x matrixNum channels = (33, 4, 4)
flankingBaseNum = 16
matrixNum = 4

# ---- batching
predictBatchSize = 1000            # sites per predict call in callVar / evaluate
trainBatchSize = 10000             # tensors per optimiser step
bloscBlockSize = 500               # rows per compressed block of a training set (utils_v2.DecompressArray)
trainingDatasetPercentage = 0.9    # first 90 % train, the rest validates

# ---- optimiser schedule (train.py)
initialLearningRate = 0.001
learningRateDecay = 0.1
maxLearningRateSwitch = 3
maxEpoch = 10000
parameterOutputPlaceHolder = 6     # digits of the epoch suffix of a checkpoint name

# ---- regularisation of the v3 networks
l2RegularizationLambda = 0.001
l2RegularizationLambdaDecay = 0.1
dropoutRateFC4 = 0.5
dropoutRateFC5 = 0.0

# ---- alignment stages (dataPrepScripts/param.py:3): reference bases fetched around a --ctgStart/--ctgEnd region
expandReferenceRegion = 1000000

# ---- TensorFlow's intra-op thread count in the reference; kept because the drivers assign it, unused here
NUM_THREADS = 12


def str2bool(v):
    """argparse type for the yes/no options (param.py:28-35): exits on anything that is not a recognised spelling"""
    s = v.lower()
    if s in ('yes', 'true', 't', 'y', '1'):
        return True
    if s in ('no', 'false', 'f', 'n', '0'):
        return False
    import sys
    raise sys.exit('Boolean value expected.')
