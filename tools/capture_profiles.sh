#!/bin/bash
# Run ON THE GPU BOX (gpurun -- 'bash tools/capture_profiles.sh <tag>'): the ncu evidence kept under profiles/.
#   1. launch list (gpu__time_duration.sum) of the default bench command, 3 steps
#   2. ncu --set full of one forward + backward of every loss variant (tools/loss_probe.py) and of the Track-W fused plan
# Each profiled command first runs once WITHOUT ncu and must exit 0.  Raw pages of the reports land in gpurun_out/<tag>_full_*.csv; turn them
# into the tracked text summaries with tools/ncu_summary.py and tools/update_traffic.py in the build container.
tag=${1:-cap}
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 3 --warmup 3 --train-steps 0 --no-cpu-baseline --train-reference-eager 0 --e2e-steps 1 --extra-configs 0"
$BENCH > $out/${tag}_plain_bench.log 2>&1 || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps3.csv $BENCH > $out/${tag}_ncu_bench.log 2>&1
python tools/loss_probe.py 2 > $out/${tag}_plain_probe.log 2>&1 || { echo "probe failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'gram|apply' --launch-skip 0 -c 16 -f -o $out/${tag}_full_loss \
    python tools/loss_probe.py 2 > $out/${tag}_ncu_probe.log 2>&1
WAVE="python bench.py --track wavelet --steps 2 --warmup 2 --no-cpu-baseline"
$WAVE > $out/${tag}_plain_wavelet.log 2>&1 || { echo "wavelet bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_wavelet.csv $WAVE > $out/${tag}_ncu_wavelet_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wavelet_res|db2_' --launch-skip 9 -c 3 -f -o $out/${tag}_full_wavelet \
    $WAVE > $out/${tag}_ncu_wavelet.log 2>&1
# the reports are large (gpurun merges at most 64 MiB back): keep their raw pages as CSV (what tools/ncu_summary.py and
# tools/update_traffic.py read) and drop the .ncu-rep files
for f in $out/${tag}_full_loss $out/${tag}_full_wavelet; do ncu -i $f.ncu-rep --page raw --csv > $f.csv 2>/dev/null && rm -f $f.ncu-rep; done
echo capture done
