mkdir -p gpurun_out
timeout 300 python scripts/bench_layers_r2.py 8 > gpurun_out/r2_conv_layers_b8.txt 2>&1; tail -75 gpurun_out/r2_conv_layers_b8.txt
MG_PDL=0 MG_TWO_STREAMS=0 timeout 300 python scripts/profile_graph_step.py 2>&1 | grep -v Warn | grep "==\| ms " > gpurun_out/r2_graph_step_kernels_serial.txt; head -32 gpurun_out/r2_graph_step_kernels_serial.txt
