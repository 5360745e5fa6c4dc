"""Model-desc strings of the BASELINE.json configurations, as the reference's recipes build them.

cfg1  README.md:52 (CIFAR10 example).
cfg2  examples/resnet34-imagenet.sh:7.
cfg3  papers/dss/denet34.sh: the ResNet-34 stack with its last three layers removed (`--layer-remove 3`, :87), skip
      sources inserted after stage 2 and stage 3 (`--layer-insert 11:SKIPSRC.X[0] 18:SKIPSRC.X[1]`, :87), batch-norm +
      relu pairs merged (`--convert-bn-relu` -> ModelCNN.convert_bn_relu()) and the 'skip' head appended (:14, :88).
"""

CIFAR_CNN = "C[128,3] BN A P[2] C[256,3] BN A P[2] C[512,3] BN A P.A R"

RESNET34 = ("C.B[64,7,2] BN A P[3,2,1] nRSN.O[3,64,3] nRSN.O[4,128,3,2] nRSN.O[6,256,3,2] nRSN.O[3,512,3,2] "
            "P.A[7] R.TB")

DENET34_SKIP_HEAD = ("PI[2] C[256,3] SKIP[1] BNA PI[2] C[128,3] SKIP[0] BNA DNC[96,100] DNS[7,24,0.01,0.1] "
                     "C[1536,1] BNA C.B[1024,1] BNA C.B[768,1] BNA C.B[512,1] BNA DND[0.5,1,1]")

DENET34_SKIP = ("C.B[64,7,2] BN A P[3,2,1] nRSN.O[3,64,3] nRSN.O[4,128,3,2] SKIPSRC.X[0] nRSN.O[6,256,3,2] "
                "SKIPSRC.X[1] nRSN.O[3,512,3,2] " + DENET34_SKIP_HEAD)

WORKLOADS = {
    # name: (model desc, data shape, per-GPU batch, classes, convert_bn_relu, solver)
    "cifar-cnn": (CIFAR_CNN, (3, 32, 32), 32, 10, False, "sgd"),
    "resnet34": (RESNET34, (3, 224, 224), 256, 1000, False, "nesterov"),
    "denet34-skip": (DENET34_SKIP, (3, 512, 512), 32, 80, True, "nesterov"),
}
