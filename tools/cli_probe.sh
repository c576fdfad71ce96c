#!/bin/bash
# Where a drop-in CLI run spends its time (SPSP_TRACE), next to the reference binaries on the same files.
set -e
t() { local label=$1; shift; local s=$(date +%s%N); "$@" > /dev/null < /dev/null; local e=$(date +%s%N); echo "$label: $(( (e - s) / 1000000 )) ms"; }
D=$(mktemp -d /dev/shm/spsp_cli_XXXX)
python - <<PY
import sys; sys.path.insert(0, ".")
from supersampler_b200 import synth
open("$D/c1.fa","wb").write(synth.fasta_bytes([("c1", synth.random_genome(5_000_000, 1))]))
g = synth.random_genome(5_000_000, 1)
open("$D/c1mut.fa","wb").write(synth.fasta_bytes([("c1mut", synth.mutate(g, 0.01, 2))]))
PY
cd $D
for i in 1 2; do
  t "ours sub_sampler -i" env SPSP_TRACE=1 $GRAFT_REPO_ROOT/supersampler_b200/bin/sub_sampler -i c1.fa -v 0
done
t "ours sub_sampler -i (c1mut)" $GRAFT_REPO_ROOT/supersampler_b200/bin/sub_sampler -i c1mut.fa -v 0
printf "subsampled_c1.gz\nsubsampled_c1mut.gz\n" > sk.txt
t "ours comparator" env SPSP_TRACE=1 $GRAFT_REPO_ROOT/supersampler_b200/bin/comparator -f sk.txt -o ours
if [ -x $GRAFT_REPO_ROOT/oracle/_ref/sub_sampler ]; then
  mkdir ref; cd ref
  t "reference sub_sampler -i" $GRAFT_REPO_ROOT/oracle/_ref/sub_sampler -i ../c1.fa -v 0
  t "reference sub_sampler -i (c1mut)" $GRAFT_REPO_ROOT/oracle/_ref/sub_sampler -i ../c1mut.fa -v 0
  printf "subsampled_c1.gz\nsubsampled_c1mut.gz\n" > sk.txt
  t "reference comparator" $GRAFT_REPO_ROOT/oracle/_ref/comparator -f sk.txt -o ref
  cd ..
  cmp <(zcat subsampled_c1.gz) <(zcat ref/subsampled_c1.gz) && echo "sketch identical"
  cmp <(zcat ours_jaccard.csv.gz) <(zcat ref/ref_jaccard.csv.gz) && echo "jaccard identical"
fi
rm -rf $D
