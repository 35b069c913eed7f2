python -m pytest tests -x -q -m gpu -k "conv3x3 or model or bnrelu" 2>&1 | tail -3
python tools/conv_bench.py --layers 128,64,512 2>&1 | tail -3
python bench.py --no-cpu-baseline --kernel-table gpurun_out/kernels_r2b.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2b_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline_hbm']['achieved']), round(d['roofline_tensor_all']['achieved']))
d=json.load(open('gpurun_out/kernels_r2b.json'))
for k in ('unetca_conv3x3_fwd_kw','unetca_conv3x3_fwd_paired','unetca_first_pairs_fwd','unetca_convT2x2_fwd'):
    v=d[k]; print(k, round(v['ms_per_step'],3), round(v['flops']/v['ms']/1e9))
P
