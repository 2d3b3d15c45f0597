#!/bin/bash
# Refresh of the round-2 ncu evidence after the stride-2 parity-halo change (run on the GPU box): a --set full capture with source
# of the first convolution that uses it (32->64 3x3/2 @608, the 1st k_conv_tc launch) and the per-layer summary of one forward.
cd "$(dirname "$0")/.."
export YB_NO_GRAPH=1
ncu --set full --clock-control none --import-source on -f -k regex:k_conv_tc --launch-skip 0 --launch-count 1 \
    -o gpurun_out/r02_conv_tc_s2halo_L1_32to64_608x608_b16 python tools/run_forward.py > /dev/null 2>&1
ncu --set full --clock-control none -f -k regex:"k_conv_tc|k_stem_tc" -o gpurun_out/r02_yolov3_608_b16_all_convs_final python tools/run_forward.py > /dev/null 2>&1
python tools/ncu_kernels.py gpurun_out/r02_yolov3_608_b16_all_convs_final.ncu-rep > gpurun_out/r02_yolov3_608_b16_all_convs_final_summary.txt 2>&1
rm -f gpurun_out/r02_yolov3_608_b16_all_convs_final.ncu-rep
ls -la gpurun_out/r02_*final* gpurun_out/r02_conv_tc_s2halo*
