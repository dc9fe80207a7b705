#!/bin/bash
# Final-state captures for profiles/ (run under gpurun, one GPU):  tools/capture_final.sh <outdir under gpurun_out>
#  1. bench.py (events, no profiler)                      -> bench_final.json
#  2. ncu launch list of one bench step                   -> launches.csv
#  3. ncu --set full of the last step's k_accum_entries / k_ntt_pass launches (skip counts taken from the launch list)
#     -> accum.csv / ntt.csv (raw pages) for tools/traffic_from_ncu.py
set -u
OUT=gpurun_out/${1:-final}
mkdir -p "$OUT"
python bench.py > "$OUT/bench_final.json" 2> "$OUT/bench_final.err"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" $BENCH > "$OUT/bench_under_ncu.json" 2> "$OUT/l.err"
read NACC NNTT <<< "$(python - "$OUT/launches.csv" <<'PY'
import csv, sys
hdr, names = None, []
for r in csv.reader(open(sys.argv[1])):
    if 'Kernel Name' in r:
        hdr = r
    elif hdr and len(r) == len(hdr):
        names.append(r[hdr.index('Kernel Name')])
print(sum('k_accum_entries' in n for n in names), sum('k_ntt_pass' in n for n in names))
PY
)"
echo "launches: k_accum_entries $NACC k_ntt_pass $NNTT" | tee "$OUT/counts.txt"
ncu --set full --clock-control none -k regex:k_accum_entries -s $((NACC - 7)) -c 7 -o "$OUT/accum" -f $BENCH > /dev/null 2> "$OUT/a.err"
ncu -i "$OUT/accum.ncu-rep" --page raw --csv > "$OUT/accum.csv" 2>> "$OUT/a.err"
ncu --set full --clock-control none -k regex:k_ntt_pass -s $((NNTT - 26)) -c 26 -o "$OUT/ntt" -f $BENCH > /dev/null 2> "$OUT/n.err"
ncu -i "$OUT/ntt.ncu-rep" --page raw --csv > "$OUT/ntt.csv" 2>> "$OUT/n.err"
rm -f "$OUT/accum.ncu-rep" "$OUT/ntt.ncu-rep"
ls -la "$OUT"
