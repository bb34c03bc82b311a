"""Submodule invocator: `python -m clairvoyante_b200 SubmoduleName [options]`; counterpart of the reference's top-level
clairvoyante.py (:1-48), which dispatches `clairvoyante.py callVarBam ...` to the script of that name."""
import importlib
import sys

SUBMODULES = ["callVarBamParallel", "callVarBam", "callVar", "calTrainDevDiff", "evaluate", "tensor2Bin", "trainNonstop", "train",
              "trainWithoutValidationNonstop", "CreateTensor", "ExtractVariantCandidates", "GetTruth", "PairWithNonVariants"]
NOT_PORTED = ["demoRun", "evaluateListOfModels", "getEmbedding", "getTensorAndLayerPNG", "ChooseItemInBed",
              "CombineMultipleDatasetsForTraining", "CountNumInBed", "RandomSampling"]     # outside the hot path


def main():
    if len(sys.argv) <= 1 or sys.argv[1] in ("-h", "--help"):
        print("clairvoyante_b200 submodule invocator:")
        print("  Usage: python -m clairvoyante_b200 SubmoduleName [Options of the submodule]")
        print("")
        print("Available submodules:")
        for n in SUBMODULES:
            print("  - %s" % n)
        sys.exit(0)
    name = sys.argv[1]
    if name not in SUBMODULES:
        sys.exit("%s is %s" % (name, "a reference submodule outside the scope of this package" if name in NOT_PORTED else "not a submodule"))
    mod = importlib.import_module("clairvoyante_b200." + name)
    sys.argv = [name + ".py"] + sys.argv[2:]
    mod.main()


if __name__ == "__main__":
    main()
