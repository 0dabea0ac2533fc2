#!/bin/bash
# Recipe: vendor the UNMODIFIED reference's Python sources into oracle/_ref/ so that they travel to the GPU box
# (oracle/_ref/ is git-ignored - reference sources never enter the history - but NOT gpurun-ignored).
# Used there by bench.py --impl reference (the reference's own train_epoch_with_grad_clip on the host cores),
# bench.py's reference_cuda leg, and scripts/acceptance_run.py (the reference's evaluate_all_metrics).
# Only runs where /root/reference exists (the build container); elsewhere the prebuilt copy is used as is.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${WGG_REFERENCE_ROOT:-/root/reference}"
if [ ! -d "$REF/src/gan" ]; then
  echo "build_ref: $REF not present, keeping oracle/_ref as is"; exit 0
fi
OUT="$HERE/_ref"
rm -rf "$OUT"
mkdir -p "$OUT/src/gan" "$OUT/src/shared" "$OUT/dataset"
cp "$REF/src/__init__.py" "$OUT/src/"
cp "$REF"/src/gan/*.py "$OUT/src/gan/"
cp "$REF"/src/shared/*.py "$OUT/src/shared/"
# src/__init__.py imports the contrastive package too
if [ -d "$REF/src/contrastive" ]; then mkdir -p "$OUT/src/contrastive" && cp "$REF"/src/contrastive/*.py "$OUT/src/contrastive/"; fi
cp "$REF/dataset/wordfreq.txt" "$OUT/dataset/"
( cd "$REF" && find src dataset/wordfreq.txt -type f -name '*.py' -o -name 'wordfreq.txt' | sort | xargs sha256sum ) > "$OUT/SHA256SUMS"
echo "build_ref: vendored $(find "$OUT" -name '*.py' | wc -l) reference files into $OUT"
