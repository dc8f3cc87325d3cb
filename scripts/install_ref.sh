#!/bin/bash
# Installs the UNMODIFIED reference (uma-pi1/kge, /root/reference) into baseline/_ref so that it travels to
# the GPU box with the gpurun snapshot (baseline/_ref is git-ignored, not gpurun-ignored).  Used by
#   * bench.py --impl reference      (the reference's own TrainingJob1vsAll on the host cores)
#   * tests/test_gpu_jobs.py         (unmodified reference jobs with `model: b200_<m>` on job.device=cuda)
# Nothing from the reference enters the git history.
#
# Step 1 is the prescribed offline pip install.  The reference's setup.py declares packages=["kge"] only (it is
# meant to be installed with `pip install -e .`), so the wheel holds the top-level modules but neither the
# sub-packages (kge.job, kge.model, kge.model.embedder, kge.util) nor the yaml package data; step 2 completes
# the SAME tree from the same source, file for file.
set -eu
cd "$(dirname "$0")/.."
REF="${KGE_REFERENCE_SRC:-/root/reference}"
[ -d "$REF/kge" ] || { echo "reference tree not found at $REF" >&2; exit 1; }
rm -rf baseline/_ref /tmp/_kge_refcopy
mkdir -p baseline
cp -r "$REF" /tmp/_kge_refcopy            # /root/reference is read-only; the build writes egg-info
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
       --target baseline/_ref /tmp/_kge_refcopy >/dev/null
# step 2: sub-packages + package data the wheel leaves out
(cd "$REF" && find kge -type f \( -name '*.py' -o -name '*.yaml' \) -print0) | \
  while IFS= read -r -d '' f; do
    mkdir -p "baseline/_ref/$(dirname "$f")"
    cp "$REF/$f" "baseline/_ref/$f"
  done
rm -rf /tmp/_kge_refcopy
find baseline/_ref -name '__pycache__' -type d -prune -exec rm -rf {} +
n_py=$(find baseline/_ref/kge -name '*.py' | wc -l); n_ref=$(find "$REF/kge" -name '*.py' | wc -l)
[ "$n_py" = "$n_ref" ] || { echo "incomplete install: $n_py of $n_ref modules" >&2; exit 1; }
echo "reference installed into baseline/_ref ($n_py modules, $(find baseline/_ref/kge -name '*.yaml' | wc -l) yaml files)"
