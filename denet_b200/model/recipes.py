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

# cfg5  papers/dss/denet101.sh "wide": ResNet-101 (bottleneck blocks [3,4,23,3], models/imagenet/resnet101) with its last
#       three layers removed, skip sources after stages 1, 2 and 3 (:90 `--layer-insert 7:SKIPSRC[0] 12:SKIPSRC.X[1]
#       24:SPLIT 37:SKIPSRC.X[2]`; SPLIT is a 12 GB-GPU memory workaround and parses as identity here) and the wide head
#       (:19) that upsamples to stride 4 (128 x 128 corner map, 48 x 48 = 2304 RoIs per image).
DENET101_WIDE_HEAD = ("PI[2] C[1024,3] SKIP[2] BNA PI[2] C[512,3] SKIP[1] BNA PI[2] C[256,3] SKIP[0] BNA SPLIT "
                      "DNC[128,200] DNS[7,48,0.01,0.1] C.B[2048,1] BNA C.B[1536,1] BNA C.B[1024,1] BNA C.B[768,1] BNA "
                      "DND[0.5,1,1]")

DENET101_WIDE = ("C.B[64,7,2] BN A P[3,2,1] nRSN.O[3,256,3,1,64] SKIPSRC[0] nRSN.O[4,512,3,2,128] SKIPSRC.X[1] "
                 "nRSN.O[23,1024,3,2,256] SKIPSRC.X[2] nRSN.O[3,2048,3,2,512] " + DENET101_WIDE_HEAD)

WORKLOADS = {
    # name: (model desc, data shape, per-GPU batch, classes, convert_bn_relu, solver)
    "cifar-cnn": (CIFAR_CNN, (3, 32, 32), 32, 10, False, "sgd"),
    "resnet34": (RESNET34, (3, 224, 224), 256, 1000, False, "nesterov"),
    "denet34-skip": (DENET34_SKIP, (3, 512, 512), 32, 80, True, "nesterov"),
    "denet101-wide": (DENET101_WIDE, (3, 512, 512), 8, 80, True, "nesterov"),
}
