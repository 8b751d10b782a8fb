export CUDA_LAUNCH_BLOCKING=1
for d in 3 1 2 0; do echo "DBG=$d"; LKF_LK_TMA_DBG=$d timeout 120 python tools/_dbg_tma.py 2>&1 | tail -2; done
LKF_LK_TMA=0 timeout 120 python tools/_dbg_tma.py 2>&1 | tail -1
