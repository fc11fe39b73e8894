#!/bin/bash
# prints what the GPU box looks like (cores, memory, first-touch speed, GPU)
echo "nproc=$(nproc)"; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader
python - <<'PY'
import time, numpy as np
for i in range(2):
    t0=time.time(); a=np.empty(1_000_000_000,dtype=np.uint8); a[:]=1; print('first touch 1GB', round(time.time()-t0,2),'s'); del a
PY
