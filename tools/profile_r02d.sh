#!/bin/bash
# ncu evidence of the final round-2 library for the kernels that changed last (weight images by TMA bulk copy) and the
# launch lists of both models (run on the GPU box through gpurun; a reduced tools/profile_round.sh).
R=r02d
set -x
cd ${GRAFT_REPO_ROOT:-.}
timeout 120 python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/${R}_quick.json 2> gpurun_out/${R}_quick.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv python bench.py --quick --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
for K in k_egno_node_fwd k_egno_node_bwd k_egno_pair; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -f -o gpurun_out/${R}_$K python bench.py --quick --no-graph --steps 1 --warmup 3 > gpurun_out/ncu_$K.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_segno.csv python tools/profile_shapes.py segno 2 > gpurun_out/ncu_ls.log 2>&1
mkdir -p gpurun_out/digests
python tools/launch_digest.py gpurun_out/${R}_launches.csv 7 > gpurun_out/digests/${R}_launches_digest.txt 2>&1
python tools/launch_digest.py gpurun_out/${R}_launches_segno.csv 2 > gpurun_out/digests/${R}_launches_segno_digest.txt 2>&1
rm -f gpurun_out/digests/ncu_digest.json
for f in gpurun_out/${R}_k_*.ncu-rep; do
  n=$(basename $f .ncu-rep)
  python tools/ncu_digest.py $f 40 > gpurun_out/digests/${n}_digest.txt 2>&1
  python tools/ncu_digest.py $f --json ${n#${R}_} profiles/${n}_digest.txt "$n" gpurun_out/digests/ncu_digest.json
  rm -f $f
done
ls -la gpurun_out/digests
