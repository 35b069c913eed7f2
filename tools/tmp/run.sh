python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "pool_backward_stream or on_the_fly" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu > gpurun_out/r1x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1x_tests.log; tail -3 gpurun_out/r1x_tests.log
UNETCA_EW_STREAM=0 python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r1x_reg.json > gpurun_out/r1x_bench_reg.json 2> gpurun_out/r1x_bench.err
python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r1x.json > gpurun_out/r1x_bench_stream.json 2>> gpurun_out/r1x_bench.err
for f in gpurun_out/r1x_bench_*.json; do python - "$f" <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline_hbm']['achieved']), round(d['roofline_tensor_all']['achieved']))
P
done
tail -3 gpurun_out/r1x_bench.err
