set -x
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -5
python tools/run_forward.py --list --reps 3 > gpurun_out/r29_layers_halo2.txt 2>&1
YB_TC_S2_HALO_MAXC=0 python tools/run_forward.py --list --reps 3 > gpurun_out/r29_layers_pertap.txt 2>&1
python tools/ab_layers.py yolov3 608 gpurun_out/r29_layers_pertap.txt gpurun_out/r29_layers_halo2.txt | grep -E "s2|total|<--"
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "pertap:YB_TC_S2_HALO_MAXC=0" "halo2:" "pertap:YB_TC_S2_HALO_MAXC=0" "halo2:"
YB_TC_STATS=1 python tools/run_forward.py 2>&1 | grep "TCSTATS.*s2" | head -6
