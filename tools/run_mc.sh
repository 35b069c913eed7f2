timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv3x3" > gpurun_out/t18_conv.log 2>&1; tail -3 gpurun_out/t18_conv.log
timeout 300 python tools/conv_bench.py --what fwd --dgrad > gpurun_out/convbench_hpix.log 2>&1; cat gpurun_out/convbench_hpix.log
timeout 300 python tools/conv_bench.py --what fwd --dgrad --no-halo > gpurun_out/convbench_nohpix.log 2>&1; cat gpurun_out/convbench_nohpix.log
