#!/bin/bash
# sweep the TMA x-chunk length on the two reference shapes
for xc in 2 3 4 6 8 10 12; do
  FDTDX_B200_TMA_XCHUNK=$xc python scripts/prof_one.py --steps 30 2>&1 | grep shape | sed "s/^/xc=$xc /"
  FDTDX_B200_TMA_XCHUNK=$xc python scripts/prof_one.py --steps 30 --shape 1897,291,128 --thickness 12 --nonuniform 2>&1 | grep shape | sed "s/^/xc=$xc /"
done
