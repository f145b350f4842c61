#!/bin/bash
# The reference's distributed mode (README.md:129-213: `sts -m i -j K` on every host, then `sts -m a -d DIR`)
# on one multi-GPU box: one stsb200 process per GPU, job K = GPU K reads its own slice of the data file
# (byte offset K * iterations * n / 8, utilities.c:1526) and writes DIR/sts.000K.<iterations>.<n>.pvalues;
# then one `stsb200 -m a -d DIR` streams all of them through the GPU tallies (bounded memory) into
# DIR/assess/result.txt - the unmodified reference's `sts -m a -d DIR` reads the same files and, when it is
# present, is run as well for comparison.  No data-path communication at all.
#
# usage: tools/run_distributed.sh <gpus> <iterations per GPU> <datafile> <workdir> [extra sts flags...]
set -euo pipefail
G=$1; I=$2; DATA=$3; DIR=$4; shift 4
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$DIR"
pids=()
for ((k = 0; k < G; k++)); do
	"$ROOT/sts_b200/host/stsb200" -m i -D "$k" -j "$k" -i "$I" -F r -w "$DIR" "$@" "$DATA" &
	pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
ls -l "$DIR"/*.pvalues
"$ROOT/sts_b200/host/stsb200" -m a -d "$DIR" -i $((G * I)) -w "$DIR/assess" "$@"
echo "assessment: $DIR/assess/result.txt"
if [ -x "$ROOT/oracle/_ref/sts" ]; then
	"$ROOT/oracle/_ref/sts" -v 0 -m a -d "$DIR" -i $((G * I)) -w "$DIR/assess_ref" "$@"
	cmp "$DIR/assess/result.txt" "$DIR/assess_ref/result.txt" && echo "identical to the reference's assessment"
fi
