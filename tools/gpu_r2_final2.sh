#!/bin/bash
# round 2, final 1-GPU call of the last session: smoke, full GPU suite, the headline bench line with every leg, the other
# BASELINE configs (device legs only: their CPU / eager columns are those of profiles/r02s_*), length-aware mode,
# launch list + DRAM traffic of one step, ncu --set full of the three dominant kernels (batch 64)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r3a_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r3a_smoke.log)"
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r3a_tests.log)"
line() { python - "$1" "$2" <<'PY'
import json, sys
wl, rc = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/r3a_{wl}.json"))
    cb, ge = d.get("cpu_baseline"), d.get("gpu_eager_baseline")
    print(f"{wl} rc={rc}: ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "gen ms", round(d["roofline"]["ms"], 2),
          "frac", round(d["roofline"]["frac"], 3), "| cpu", cb and (round(cb["value"], 2), cb["kind"], cb["cores"]), "| eager", ge and round(ge.get("value", 0), 1),
          "| dur", d.get("duration_check"), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| launches/step", d.get("gpu_launches_per_step"))
except Exception as e:
    print(wl, "rc=" + rc, "failed", e); print(open(f"gpurun_out/r3a_{wl}.err").read()[-1200:])
PY
}
wl=multilingual_v3_b256x128
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r3a_$wl.json 2> gpurun_out/r3a_$wl.err; line $wl $?
for wl in baker_v1_cli_b1 baker_v1_gen_b64x640 baker_v3_gen_b64x640 aishell3_v1_b32x512 multilingual_v3_b1024x128 baker_v1_b64x128; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl > gpurun_out/r3a_$wl.json 2> gpurun_out/r3a_$wl.err; line $wl $?
done
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --length-aware 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('length-aware: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r3a_dram_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r3a_ncu.log 2>&1; echo "ncu launch list rc=$?"
python tools/launches.py gpurun_out/r3a_dram_traffic.csv 2>&1 | head -24
python tools/traffic_from_ncu.py gpurun_out/r3a_dram_traffic.csv --quiet
prof() { # name, kernel regex, skip
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -o gpurun_out/r3a_$1 -f \
     python bench.py --steps 1 --warmup 2 --no-cpu --batch 64 --profile-range > gpurun_out/r3a_ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -n 1 gpurun_out/r3a_ncu_$1.log | cut -c1-160)"
}
prof mrf16_c32 "fused_mrf16_kernel<32" 0
prof mrf16_c64 "fused_mrf16_kernel<64" 0
prof tc16r_flow_in "conv1d_tc16r_kernel<3" 4
ls -la gpurun_out/*.ncu-rep
