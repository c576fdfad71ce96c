#!/bin/bash
# bin/sub_sampler -f + bin/comparator on the C2 files (64 x 5 Mbp on tmpfs) with SPSP_TRACE, twice
set -e
D=$(mktemp -d /dev/shm/spsp_clif_XXXX)
python - <<PY
import sys; sys.path.insert(0, ".")
from supersampler_b200 import synth
fam = synth.Family(5_000_000, 42)
names = []
for i in range(64):
    p = "$D/g%05d.fa" % i
    open(p, "wb").write(fam.fasta(i)); names.append(p)
open("$D/in.txt", "w").write("\n".join(names) + "\n")
PY
cd $D
for i in 1 2; do
  s=$(date +%s%N)
  SPSP_TRACE=1 $GRAFT_REPO_ROOT/supersampler_b200/bin/sub_sampler -f in.txt -t 16 -v 0 > /dev/null
  e=$(date +%s%N); echo "sub_sampler -f: $(( (e - s) / 1000000 )) ms"
done
ls subsampled_g*.gz > sk.txt
s=$(date +%s%N)
SPSP_TRACE=1 $GRAFT_REPO_ROOT/supersampler_b200/bin/comparator -f sk.txt -o res > /dev/null
e=$(date +%s%N); echo "comparator: $(( (e - s) / 1000000 )) ms"
rm -rf $D
