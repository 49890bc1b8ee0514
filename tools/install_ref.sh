#!/bin/bash
# Reference arm of bench.py (`--impl reference`): the reference is a collection of Python scripts without setup.py /
# pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to build.  "Installing" it means
# placing its UNMODIFIED .py files under baseline/_ref/ (git-ignored: never part of this repo's history; NOT
# gpurun-ignored: it travels to the GPU box with the snapshot), where bench.py imports train.py from.
#   tools/install_ref.sh [reference checkout, default /root/reference]
set -e
SRC=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
if [ ! -f "$SRC/train.py" ]; then
  echo "install_ref: no reference checkout at $SRC (nothing installed; bench.py --impl reference falls back to the oracle port)" >&2
  exit 0
fi
mkdir -p "$ROOT/baseline/_ref"
cp "$SRC"/*.py "$ROOT/baseline/_ref/"
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$ROOT/baseline/_ref/REFERENCE_COMMIT"
( cd "$ROOT/baseline/_ref" && sha256sum *.py ) > "$ROOT/baseline/_ref/SHA256SUMS"
echo "install_ref: $(ls "$ROOT/baseline/_ref"/*.py | wc -l) reference files -> $ROOT/baseline/_ref"
