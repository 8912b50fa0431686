#!/bin/bash
# 8-GPU: BASELINE config 5 (65536^2, rain + evaporation, open boundary, fp64 mass ledger) with the band kernel and the fused exchange.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 scripts/config5_ledger.py --size 65536 --steps 200 --fast-steps 2000 2>gpurun_out/config5.err | grep CONFIG5 > gpurun_out/config5.json
cat gpurun_out/config5.json; tail -3 gpurun_out/config5.err
