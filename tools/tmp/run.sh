python -m pytest tests -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log; tail -3 gpurun_out/r2a_tests.log
python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r2a.json > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2a_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline_hbm']['achieved']), round(d['roofline_tensor_all']['achieved']))
d=json.load(open('gpurun_out/kernels_r2a.json'))
for k in ('unetca_bn_relu',):
    v=d[k]; print(k, round(v['ms_per_step'],3), round(v['bytes']/v['ms']/1e6))
P
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3_kw -s 6 -c 1 -f -o gpurun_out/r2a_kw $CMD > gpurun_out/r2a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3_pixn -s 15 -c 1 -f -o gpurun_out/r2a_pixn2 $CMD > gpurun_out/r2a_ncu2.log 2>&1
for r in r2a_kw r2a_pixn2; do ncu -i gpurun_out/$r.ncu-rep --page details > gpurun_out/$r.details.txt 2>&1; done
ls -la gpurun_out/r2a_*
