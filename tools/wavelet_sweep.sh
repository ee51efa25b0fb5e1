#!/bin/bash
# Track W sweep (BASELINE configs[4] shape family): Haar vs db2, J = 1..5, 512 and 1024 maps, fused plan vs per-level kernels.
# Prints one line per configuration; run on a B200 box from the repo root.
for size in 512 1024; do
  for wv in haar db2; do
    for J in 1 2 3 4 5; do
      for res in 1 0; do
        python bench.py --track wavelet --size $size --wavelet-name $wv --wavelet-levels $J --wavelet-resident $res --steps 50 --warmup 5 2>/dev/null | tail -1 | \
          python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$size $wv J=$J', 'fused' if $res else 'per-level', '| ms %.4f | Gpix/s %.1f | frac(8 B/elem) %.3f | %s' % (d['ms_per_step'], d['value']/1e3, d['roofline']['frac'], d['config']['path']))"
      done
    done
  done
done
