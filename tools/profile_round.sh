#!/bin/bash
# ncu evidence of a round (run on the GPU box through gpurun): plain run first, then the launch list, then one
# --set full capture per edge kernel; digests are made here with tools/ncu_digest.py / tools/launch_digest.py
set -x
cd ${GRAFT_REPO_ROOT:-.}
timeout 120 python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err || exit 1
cat gpurun_out/quick.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01b_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_bwd_sel -s 8 -c 1 -f -o gpurun_out/r01b_edge_bwd python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_b.log 2>&1
if [ -z "$SKIP_FWD" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_fwd_sel -s 8 -c 1 -f -o gpurun_out/r01b_edge_fwd python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_f.log 2>&1
fi
ls -la gpurun_out/
