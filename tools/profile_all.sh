#!/bin/bash
# Round profile capture (run under gpurun): launch list of the bench command + one ncu --set full capture per kernel family.
# Every program is first run to completion WITHOUT ncu.
set -x
O=gpurun_out
R=r2
python bench.py --steps 3 --warmup 3 > $O/${R}_bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${R}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn1_sweep -c 1 -o $O/${R}_prof_sweep -f python bench.py --steps 3 --warmup 3 > $O/${R}_ncu_sweep.log 2>&1
python tools/nn1_one.py && ncu --set full --clock-control none --import-source on -k regex:"nn1_(arm|fixup|bwd)" -c 3 -o $O/${R}_prof_nn1_small -f python tools/nn1_one.py > $O/${R}_ncu_nn1_small.log 2>&1
python tools/knn_one.py 32 2048 64 20 && ncu --set full --clock-control none --import-source on -k regex:knnc_kernel -c 1 -o $O/${R}_prof_knnc -f python tools/knn_one.py 32 2048 64 20 > $O/${R}_ncu_knnc.log 2>&1
python tools/knn_one.py 32 2048 3 20 && ncu --set full --clock-control none --import-source on -k regex:"knn3_|knn_threshold" -c 5 -o $O/${R}_prof_knn3 -f python tools/knn_one.py 32 2048 3 20 > $O/${R}_ncu_knn3.log 2>&1
python tools/graph_one.py 32 64 2048 20 && ncu --set full --clock-control none --import-source on -k regex:"edge_feature|edge_csr|fps_kernel|clip_|offset_gather" -c 14 -o $O/${R}_prof_graph -f python tools/graph_one.py 32 64 2048 20 > $O/${R}_ncu_graph.log 2>&1
python tools/geom_one.py 32 2048 && ncu --set full --clock-control none --import-source on -k regex:"local_frames|kappa_|knn_outlier" -c 5 -o $O/${R}_prof_geom -f python tools/geom_one.py 32 2048 > $O/${R}_ncu_geom.log 2>&1
ls -la $O/${R}_prof_*.ncu-rep
