timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench_configs.py --configs 3,3L,3D 2>&1 | grep '^{' > gpurun_out/r2_configs_cfg3_packed.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2_configs_cfg3_packed.jsonl"):
    d=json.loads(l); print(d["config"], d["ms_per_pass"], d["roofline"]["frac"], d.get("attempts"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_cfg3_2p20_packed.csv python tools/prof_pass.py --config 3L > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:cnf_rk --launch-skip 1 --launch-count 1 -o gpurun_out/r2_cnf_attempt_v3 python tools/prof_pass.py --config 3L > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:cnf_rk_adj --launch-count 1 -o gpurun_out/r2_cnf_adj_v3 python tools/prof_pass.py --config 3L > /dev/null 2>&1
