#!/bin/bash
# ncu captures of round 2 (run under gpurun on ONE GPU): per-kernel time / DRAM bytes / tensor-pipe activity of the
# scene-inference step and of the fused training step, and the launch list of bench.py.  Outputs under gpurun_out/.
set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active
SCENE='conv0_tc|x16_tile|spectral_logits|conv1_pool|conv2_scene|pool2_cls|head_sum'
TRAIN='train_conv|patch_cnn|multi_gemm|sim_tc|head_fwd|head_bwd|head_wgrad|loss_rows|loss_graph|adam_all'
ncu --metrics $M --clock-control none -k regex:"$SCENE" -s 14 -c 7 --csv --log-file gpurun_out/r02_scene_kernels.csv python scripts/profile_scene.py 4 > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:"$TRAIN" -s 15 -c 15 --csv --log-file gpurun_out/r02_train_kernels.csv python scripts/profile_train.py 3 > /dev/null 2>&1
ls -la gpurun_out/r02_*
