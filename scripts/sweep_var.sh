#!/bin/bash
# usage: sweep_var.sh lib1.so lib2.so ... : standard two shapes per library variant
for lib in "$@"; do
  FDTDX_B200_LIB=$lib python scripts/prof_one.py --steps 30 2>&1 | grep shape | sed "s#^#$(basename $lib) #"
  FDTDX_B200_LIB=$lib python scripts/prof_one.py --steps 30 --shape 1897,291,128 --thickness 12 --nonuniform 2>&1 | grep shape | sed "s#^#$(basename $lib) #"
done
