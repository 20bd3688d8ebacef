#!/usr/bin/env bash
# N concurrent D2H probes, one per GPU, with and without NUMA binding (what bounds bench.py's e2e at N > 1).
N=${1:-8}
for mode in "" "--numa"; do
  T=$(python -c "import time; print(time.time() + 25)")
  for i in $(seq 0 $((N-1))); do
    python tools/pcie_probe.py --device $i --reps 3 --start-at $T $mode > gpurun_out/pcie_conc_${N}_${i}${mode}.json 2>/dev/null &
  done
  wait
  echo "mode=[$mode]"; cat gpurun_out/pcie_conc_${N}_*${mode}.json | python -c "
import sys, json
rows=[json.loads(l) for l in sys.stdin if l.startswith('{')]
print('d2h_whole_gbs per GPU:', [r['d2h_whole_gbs'] for r in rows], 'sum', round(sum(r['d2h_whole_gbs'] for r in rows),1))
print('generate_batch_host_gbs per GPU:', [r['generate_batch_host_gbs'] for r in rows], 'sum', round(sum(r['generate_batch_host_gbs'] for r in rows),1))
print('numa:', [r['numa_node'] for r in rows])"
  rm -f gpurun_out/pcie_conc_${N}_*${mode}.json
done
nproc; ls /sys/devices/system/node/ | grep node
