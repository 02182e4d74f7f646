#!/bin/bash
# ncu evidence of a round (run on the GPU box through gpurun): plain run first, then the launch list, then one
# --set full capture per dominant kernel; digests are made in the build container with tools/ncu_digest.py /
# tools/launch_digest.py and copied to profiles/ (tools/make_digests.sh).
R=${ROUND_TAG:-r02}
set -x
cd ${GRAFT_REPO_ROOT:-.}
timeout 120 python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/${R}_quick.json 2> gpurun_out/${R}_quick.err || exit 1
cat gpurun_out/${R}_quick.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
for K in k_edge_bwd_sel k_edge_fwd_sel k_egno_node_fwd k_egno_node_bwd k_gemm64_tc k_wgrad64_tc k_tconv_bwd; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o gpurun_out/${R}_$K python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_$K.log 2>&1
done
# SEGNO (configs[3] shape) and the blocked walk (configs[4] shape): tools/profile_shapes.py runs one training step of each
for K in k_segno_fused_fwd k_segno_node_bwd; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f -o gpurun_out/${R}_$K python tools/profile_shapes.py segno > gpurun_out/ncu_$K.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_segno.csv python tools/profile_shapes.py segno 2 > gpurun_out/ncu_ls.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_edge_bwd_sel -s 8 -c 1 -f -o gpurun_out/${R}_k_edge_bwd_sel_segno python tools/profile_shapes.py segno > gpurun_out/ncu_sbwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_edge_bwd_sel -s 2 -c 1 -f -o gpurun_out/${R}_k_edge_bwd_sel_blk python tools/profile_shapes.py egno100 > gpurun_out/ncu_bblk.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_edge_fwd_sel -s 2 -c 1 -f -o gpurun_out/${R}_k_edge_fwd_sel_blk python tools/profile_shapes.py egno100 > gpurun_out/ncu_fblk.log 2>&1
# digests are made here (the reports are too large to travel back: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out/digests
python tools/launch_digest.py gpurun_out/${R}_launches.csv 7 > gpurun_out/digests/${R}_launches_digest.txt 2>&1
python tools/launch_digest.py gpurun_out/${R}_launches_segno.csv 2 > gpurun_out/digests/${R}_launches_segno_digest.txt 2>&1
for f in gpurun_out/${R}_k_*.ncu-rep; do
  n=$(basename $f .ncu-rep)
  python tools/ncu_digest.py $f 40 > gpurun_out/digests/${n}_digest.txt 2>&1
  python tools/ncu_digest.py $f --json ${n#${R}_} profiles/${n}_digest.txt "$n" gpurun_out/digests/ncu_digest.json
  rm -f $f
done
ls -la gpurun_out/digests
