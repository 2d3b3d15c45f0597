timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -4
python tools/run_forward.py --list --reps 5 > gpurun_out/r36_alt.txt 2>&1
YB_TC_NO_EPI_ALT_S2=1 python tools/run_forward.py --list --reps 5 > gpurun_out/r36_noalt.txt 2>&1
echo "--- no alt vs alt"; python tools/ab_layers.py yolov3 608 gpurun_out/r36_noalt.txt gpurun_out/r36_alt.txt | grep -E "s2|total"
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "noalt:YB_TC_NO_EPI_ALT_S2=1" "alt:" "noalt:YB_TC_NO_EPI_ALT_S2=1" "alt:"
YB_TC_STATS=1 python tools/run_forward.py 2>&1 | grep "TCSTATS" | head -2
