python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2i_bench.json')); print(round(d['value'],1), round(d['ms_per_step'],2), d['e2e'], d['final_loss'])
P
tail -3 gpurun_out/r2i_bench.err
