#!/bin/bash
# One-GPU evidence run (under gpurun): GPU tests, the bench line, the reference arm, the ncu launch list, one
# full-metric capture of two frames and a seeded GPU fuzz.  Everything lands in gpurun_out/<tag>_*; tools/ncu_summary.py
# turns the report into the text files of profiles/ afterwards (in the build container).
#     gpurun --timeout 1500 -- 'bash tools/evidence.sh r02'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --pipeline-depth 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fragments|k_dof|k_spans|k_setup|k_vertex|k_mark" -s 24 -c 12 \
    -f -o $out/${tag}_prof $B > $out/${tag}_ncu.log 2>&1
python tools/fuzz_gpu.py 1000 1300 $out/${tag}_fuzz_gpu.json > $out/${tag}_fuzz.log 2>&1
cat $out/${tag}_tests.log
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "serial", d["one_frame_at_a_time"]["fps"], "e2e", d["e2e"]["value"], d["e2e"]["blocking_call_fps"], "fnv", d["frame_fnv_ok"])
print({k: round(v, 4) for k, v in d["ms_per_stage"].items()})
print([(k["kernel"], round(k["frac"], 3)) for k in d["roofline"]["all_kernels"]], d["cpu_baseline"]["value"])
print("also", d["also"]["fps"], d["also"]["one_frame_at_a_time_fps"], d["also"]["e2e_fps"], "sharded", d["sharded_frame"]["ms_per_frame"], "multiview", d["multiview_frame"]["ms_per_frame"])
print("others", {k: (round(v["one_frame_at_a_time_fps"]), round(v["ms_per_frame"], 4)) for k, v in d.get("other_workloads", {}).items()}, "cpu frags/s", d["cpu_baseline"].get("fragments_shaded_per_s"))
print("ref arm", json.loads(open("$out/${tag}_ref.json").read().strip().splitlines()[-1])["value"])
PY
tail -2 $out/${tag}_fuzz.log
ls -la $out/${tag}_prof.ncu-rep
