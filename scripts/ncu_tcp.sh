#!/bin/bash
# development aid: one ncu --set full capture of the tcp reverse (and optionally forward) kernel of config 1 (run under gpurun)
N=${1:-262144}
ncu --set full --clock-control none --import-source on -k regex:tc_bwd_kernel -c 1 -o gpurun_out/tcp_bwd -f python scripts/kernel_bench.py $N > gpurun_out/ncu_bwd.log 2>&1
if [ "$2" = "fwd" ]; then
ncu --set full --clock-control none --import-source on -k regex:tcp_fwd_kernel -c 1 -o gpurun_out/tcp_fwd -f python scripts/kernel_bench.py $N > gpurun_out/ncu_fwd.log 2>&1
fi
tail -n 2 gpurun_out/ncu_bwd.log
