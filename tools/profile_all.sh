#!/bin/bash
# Round profile capture (run under gpurun): launch list of the bench command + one ncu --set full capture per kernel family.
# Every program is first run to completion WITHOUT ncu.
set -x
O=gpurun_out
python bench.py --steps 3 --warmup 3 > $O/bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn1_sweep -c 1 -o $O/prof_sweep -f python bench.py --steps 3 --warmup 3 > $O/ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"nn1_(prep|fixup|reduce|bwd)" -c 5 -o $O/prof_nn1_small -f python bench.py --steps 3 --warmup 3 > $O/ncu_nn1_small.log 2>&1
python tools/knn_one.py 32 2048 64 20 && ncu --set full --clock-control none --import-source on -k regex:knnc_kernel -c 1 -o $O/prof_knnc -f python tools/knn_one.py 32 2048 64 20 > $O/ncu_knnc.log 2>&1
python tools/knn_one.py 32 2048 3 20 && ncu --set full --clock-control none --import-source on -k regex:"knn3_|knn_threshold" -c 5 -o $O/prof_knn3 -f python tools/knn_one.py 32 2048 3 20 > $O/ncu_knn3.log 2>&1
python tools/graph_one.py 32 64 2048 20 && ncu --set full --clock-control none --import-source on -k regex:"edge_feature|fps_kernel" -c 6 -o $O/prof_graph -f python tools/graph_one.py 32 64 2048 20 > $O/ncu_graph.log 2>&1
ls -la $O/*.ncu-rep
