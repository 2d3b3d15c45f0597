timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -4
python tools/run_forward.py --list --reps 5 > gpurun_out/r37_new.txt 2>&1
YB_LIB=$PWD/yolo2_light_b200/libyb_prev.so python tools/run_forward.py --list --reps 5 > gpurun_out/r37_prev.txt 2>&1
echo "--- prev vs new"; python tools/ab_layers.py yolov3 608 gpurun_out/r37_prev.txt gpurun_out/r37_new.txt | awk '$NF=="<--" || /total/ || $6+0 > 1.03 || ($6+0 < 0.97 && $6+0 > 0)'
STEPS=30 BENCH_ARGS=--no-cpu-baseline bash tools/ab_bench.sh "prev:YB_LIB=$PWD/yolo2_light_b200/libyb_prev.so" "new:" "prev:YB_LIB=$PWD/yolo2_light_b200/libyb_prev.so" "new:"
