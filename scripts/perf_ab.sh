#!/bin/bash
# scripts/perf_ab.sh -- A/B timing of alternative builds of libskr.so (build/libskr_<name>.so, see host/Makefile SKR_DEFS).
# usage: scripts/perf_ab.sh "<configs>" name1 name2 ...
cfgs=$1; shift
for n in "$@"; do
  echo "== $n"
  SKR_LIB=$(pwd)/build/libskr_$n.so timeout 200 python scripts/perf_configs.py $cfgs 2>&1 | tail -2
done
