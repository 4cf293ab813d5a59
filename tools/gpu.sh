#!/bin/bash
# build (always from the repo root) then run a command on the GPU box
set -e
cd /root/repo
python -m cra5_b200.build | tail -1
/usr/local/graft/bin/gpurun --timeout ${GPU_TIMEOUT:-900} -- "$@"
