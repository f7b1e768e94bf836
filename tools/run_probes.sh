#!/bin/bash
# usage: run_probes.sh N bin1 bin2 ...
N=$1; shift
for b in "$@"; do echo -n "$b: "; timeout 120 ./tools/bin/$b $N 3 1 0 || echo FAIL; done
