cd .exp
./tma_probe 256 8 0 1 0 8192 0 96
./tma_probe 256 8 0 1 0 8192 0 104
./tma_probe 48 24 2 2 2688 8704 1 112
./tma_probe 48 24 2 2 2688 8704 1 116
./tma_probe 32 24 5 2 3840 8712 2 -16
