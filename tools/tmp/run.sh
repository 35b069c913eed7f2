python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "outc" 2>&1 | tail -5
python -m pytest tests -x -q -m gpu > gpurun_out/r1z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1z_tests.log; tail -3 gpurun_out/r1z_tests.log
python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r1z.json > gpurun_out/r1z_bench_stream.json 2>> gpurun_out/r1z_bench.err
for f in gpurun_out/r1z_bench_*.json; do python - "$f" <<'P'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline_hbm']['achieved']), round(d['roofline_tensor_all']['achieved']))
P
done
python - <<'P'
import json
d=json.load(open('gpurun_out/kernels_r1z.json'))
for k in ('unetca_outc_fwd','unetca_outc_bwd'):
    v=d[k]; print(k, round(v['ms_per_step'],3), round(v['bytes']/v['ms']/1e6))
P
tail -3 gpurun_out/r1z_bench.err
