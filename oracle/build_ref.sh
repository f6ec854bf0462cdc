#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference native module.
#
# Compiles the reference's own Cython shim + C++ dynamics file from where they
# lie under /root/reference (read-only) into oracle/_ref/.  Nothing from the
# reference is copied into the repository: the generated C++ lives in a temp
# dir and only the compiled extension module lands in oracle/_ref/
# (git-ignored, but it travels to the GPU box with the gpurun snapshot).
#
# Flags mirror what `python setup.py build_ext` of the reference produces
# (survey SURVEY.md section 8c: g++ -O2, no -march, hence no FMA contraction).
set -euo pipefail
REF=${DPILQR_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/dpilqr/bbdynamicswrap.pyx" ]; then
    echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
PY=${PYTHON:-python}
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
NPINC="$($PY -c 'import numpy; print(numpy.get_include())')"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
$PY -m cython -3 --cplus "$REF/dpilqr/bbdynamicswrap.pyx" -o "$TMP/bbdynamicswrap.cpp"
g++ -O2 -g0 -DNDEBUG -fPIC -shared -fwrapv -w \
    -I"$PYINC" -I"$NPINC" -I"$REF/dpilqr" \
    "$TMP/bbdynamicswrap.cpp" -o "$OUT/bbdynamicswrap$EXT"
echo "built $OUT/bbdynamicswrap$EXT"
