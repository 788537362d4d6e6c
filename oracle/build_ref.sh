#!/bin/sh
# Compiles the only piece of the reference that is C: the (unlinked) SSE prototype of the batch
# matrix * vector routine, /root/reference/c/matrix_amd64.c, straight from where it lies, into
# oracle/_ref/ (git-ignored, travels to the GPU box).  The Go renderer itself cannot be built here
# (no Go toolchain).  NOTE the prototype sums (p1+p2)+(p3+p4), not the ((p1+p2)+p3)+p4 of the Go
# assembly: it validates the oracle's products and, for vectors with z == 0
# (where both orders coincide), its results bit for bit — it is not the parity target.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$REF/c/matrix_amd64.c" ] || { echo "no reference checkout at $REF: skipping oracle/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
${CC:-gcc} -O2 -msse2 -ffp-contract=off -shared -fPIC -include stddef.h -I"$REF/c" \
    -o "$HERE/_ref/libref_cmatrix.so" "$REF/c/matrix_amd64.c"
echo "built $HERE/_ref/libref_cmatrix.so"
