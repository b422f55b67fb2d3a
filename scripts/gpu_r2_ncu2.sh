# ncu captures after the single-thread MMA issue change: weight gradient and forward convolution of D0.conv1
mkdir -p gpurun_out
python scripts/one_conv.py wgrad 16 32 512 8 > gpurun_out/plain5.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_wgrad -s 3 -c 1 -o gpurun_out/r02_prof_wgrad_d0c1 -f python scripts/one_conv.py wgrad 16 32 512 8 > gpurun_out/ncu5.log 2>&1
python scripts/one_conv.py fprop 16 32 512 8 > gpurun_out/plain6.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3 -s 3 -c 1 -o gpurun_out/r02_prof_conv_fprop_d0c1_v2 -f python scripts/one_conv.py fprop 16 32 512 8 > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3; tail -2 gpurun_out/ncu5.log gpurun_out/ncu6.log
