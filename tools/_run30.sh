timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -5
python tools/run_forward.py --list --reps 3 > gpurun_out/r30_new.txt 2>&1
YB_TC_NO_S2_TMA_EPI=1 python tools/run_forward.py --list --reps 3 > gpurun_out/r30_lsu.txt 2>&1
YB_TC_S2_CG1=1 python tools/run_forward.py --list --reps 3 > gpurun_out/r30_cg1.txt 2>&1
echo "--- lsu-epilogue vs new"; python tools/ab_layers.py yolov3 608 gpurun_out/r30_lsu.txt gpurun_out/r30_new.txt | grep -E "s2|total"
echo "--- cg1 vs new"; python tools/ab_layers.py yolov3 608 gpurun_out/r30_cg1.txt gpurun_out/r30_new.txt | grep -E "s2|total"
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "old:YB_TC_S2_HALO_MAXC=0,YB_TC_NO_S2_TMA_EPI=1" "new:" "old:YB_TC_S2_HALO_MAXC=0,YB_TC_NO_S2_TMA_EPI=1" "new:"
YB_TC_STATS=1 python tools/run_forward.py 2>&1 | grep "TCSTATS.*s2" | head -6
