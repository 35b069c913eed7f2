python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "stream" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu > gpurun_out/r1y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1y_tests.log; tail -3 gpurun_out/r1y_tests.log
python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r1y.json > gpurun_out/r1y_bench_stream.json 2>> gpurun_out/r1y_bench.err
UNETCA_EW_STREAM=0 python bench.py --no-cpu-baseline > gpurun_out/r1y_bench_reg.json 2> gpurun_out/r1y_bench.err
for f in gpurun_out/r1y_bench_*.json; do python - "$f" <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline_hbm']['achieved']), round(d['roofline_tensor_all']['achieved']))
P
done
tail -3 gpurun_out/r1y_bench.err
