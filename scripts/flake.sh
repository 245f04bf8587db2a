#!/bin/bash
# stress: the sharded host-expanded CLI run against the one-engine run, many times; usage: flake.sh <bin> <runs> [extra CLI flags]
BIN=$1; RUNS=$2; shift 2
P="gene-length=100,num-genomes=10,core-size=20,theta=200,rho=0.2,seed=11"
T=$(mktemp -d)
$BIN -o $T/one -p $P --gpu-format > /dev/null 2>&1
bad=0
for i in $(seq 1 $RUNS); do
	rm -rf $T/many; $BIN -o $T/many -p $P --devices 3 "$@" > /dev/null 2>&1
	for f in $T/one/*; do
		if ! cmp -s $f $T/many/$(basename $f); then bad=$((bad+1)); echo "run $i: $(basename $f) differs: $(cmp $f $T/many/$(basename $f) 2>&1 | head -1)"; fi
	done
done
echo "$BIN $*: $bad differing files in $RUNS runs"
rm -rf $T
