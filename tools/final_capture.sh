#!/bin/bash
# The round's evidence in one GPU call (run on the box from the repo root: bash tools/final_capture.sh <outdir under gpurun_out>):
# ncu launch list + --set full of config 2 -> summary + profiles/traffic.json (keyed by the kernel-source hash), ncu summaries of
# configs 3, 4, 5, smoke(), and the bench line (which then quotes the fresh traffic).  Everything lands in gpurun_out/<outdir>/.
set -u
O=gpurun_out/${1:-final}
mkdir -p $O
NCU="ncu --clock-control none"
if [ "${2:-all}" != "noc2" ]; then
# ---- config 2: launch list of the bench command, then one full capture per kernel class
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/c2_launches.csv python bench.py --steps 2 --warmup 3 --time 64 --no-cpu --no-e2e --no-other > $O/bench_under_ncu.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"cols_async_kernel|rowsz_power_kernel" --launch-skip 2 -c 2 -o /tmp/c2full python tools/run_one.py c2 32 2 > $O/ncu_c2.log 2>&1
ncu -i /tmp/c2full.ncu-rep --page raw --csv > $O/c2_raw.csv 2>> $O/ncu_c2.log
ncu -i /tmp/c2full.ncu-rep --page source --csv --print-source sass > $O/c2_src_sass.csv 2>> $O/ncu_c2.log
python tools/summarize_profile.py $O/c2_launches.csv $O/c2_raw.csv $O/r02_ncu_summary_config2.md $O/traffic.json 536870912 > $O/summarize.log 2>&1
cp $O/traffic.json profiles/traffic.json
fi
# ---- configs 3, 4, 5: launch list + full capture of one repetition
for c in c3both:16:67108864 c4:2048:536870912 c5:8192:268435456; do
  IFS=: read what batch pts <<< "$c"
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${what}_launches.csv python tools/run_one.py $what $batch 2 > $O/ncu_${what}.log 2>&1
  timeout 600 $NCU --set full -k regex:"^(cols|rows|rowline|pad_|permute|spectral|roll_|moments|binned|dft_|smooth|bluestein|fourstep|herm|mirror)" --launch-skip $(python - <<PY
import csv
rows=[l for l in open("$O/${what}_launches.csv") if not l.startswith("==")]
n=sum(1 for r in csv.DictReader(rows) if "xrftb" in r["Kernel Name"])
print(n//2)
PY
) -c 12 -o /tmp/${what}full python tools/run_one.py $what $batch 2 >> $O/ncu_${what}.log 2>&1
  ncu -i /tmp/${what}full.ncu-rep --page raw --csv > $O/${what}_raw.csv 2>> $O/ncu_${what}.log
done
# ---- smoke + the bench line (traffic non-null: the capture above is of these sources)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
ls -la $O
tail -2 $O/smoke.log
head -c 1500 $O/bench_n1.json
