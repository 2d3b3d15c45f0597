#!/bin/bash
# Round-2 ncu evidence, run on the GPU box (gpurun): everything lands in gpurun_out/r02_*; tools/ncu_kernels.py turns the
# reports into the text summaries that are committed under profiles/.
#   1. --set full capture (with source) of the dominant kernel: 128->256 3x3 @76x76 + shortcut (layer 14 = 11th k_conv_tc launch)
#   2. --set full captures of one launch of every other convolution class of yolov3-608 b16
#   3. the INT8 / XNOR-as-i8 / popcount convolution kernels of BASELINE configs[2] / configs[3], the fused stems
#   4. the launch list of the bench command (gpu__time_duration.sum per launch) and the DRAM traffic of every k_conv_tc launch
cd "$(dirname "$0")/.."
N="ncu --set full --clock-control none --import-source on -f"
M="ncu --set full --clock-control none -f"
export YB_NO_GRAPH=1
$N -k regex:k_conv_tc --launch-skip 10 --launch-count 1 -o gpurun_out/r02_conv_tc_cg2_halo_tmaepi_L14_128to256_76x76_b16 python tools/run_forward.py > /dev/null 2>&1
# all convolution launches of one forward, metrics only where possible: full set, no source (one report, ~75 kernels)
$M -k regex:"k_conv_tc|k_stem_tc" -o gpurun_out/r02_yolov3_608_b16_all_convs python tools/run_forward.py > /dev/null 2>&1
$M -k regex:"k_conv_tc|k_stem_pool|k_conv_stem" -o gpurun_out/r02_int8_tiny416_b64_convs python tools/run_forward.py --model yolov3-tiny --size 416 --batch 64 --quantized 1 > /dev/null 2>&1
$M -k regex:"k_conv_tc|k_conv_xnor|k_stem_pool|k_conv_stem" -o gpurun_out/r02_xnor_416_b64_convs python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 > /dev/null 2>&1
YB_XNOR_TC=0 $M -k regex:"k_conv_xnor" -o gpurun_out/r02_xnor_416_b64_popcount_only python tools/run_forward.py --model tiny-yolo-obj_xnor --size 416 --batch 64 > /dev/null 2>&1
unset YB_NO_GRAPH
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
bash tools/ncu_traffic.sh
for f in r02_yolov3_608_b16_all_convs r02_int8_tiny416_b64_convs r02_xnor_416_b64_convs r02_xnor_416_b64_popcount_only; do
  python tools/ncu_kernels.py gpurun_out/$f.ncu-rep > gpurun_out/${f}_summary.txt 2>&1
done
# gpurun merges at most 64 MiB back: keep the one report with source (the dominant kernel) and the text summaries of the rest
rm -f gpurun_out/r02_yolov3_608_b16_all_convs.ncu-rep gpurun_out/r02_int8_tiny416_b64_convs.ncu-rep gpurun_out/r02_xnor_416_b64_convs.ncu-rep \
      gpurun_out/r02_xnor_416_b64_popcount_only.ncu-rep
ls -la gpurun_out/r02_*
