#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/libturner_ref_{pathtracer,raycaster,raytracer}.so
# from the reference's own sources where they lie under /root/reference:
#   lib/kdtree.cpp + pathtracer.cpp (or raycaster.cpp) + oracle/ref_driver.cpp
# Nothing from /root/reference is copied into the repo. Because the reference's
# headers include each other by relative path ("../src/geometry.h"), the build
# uses a throw-away mirror of symlinks under $TMPDIR; the only file that is not a
# symlink there is src/geometry.h, which g++ >= 11 rejects as shipped (in-class
# friend function *template definitions* at src/geometry.h:413-415 and :740-743
# are re-defined once per class instantiation). The mirror hoists exactly those
# two templates out of their classes, bodies unchanged. The mirror is deleted
# afterwards; the only outputs are the two .so files under oracle/_ref/.
# Flags: -O2 -DNDEBUG -ffp-contract=off, no -march (the reference's CMakeLists.txt:3-7
# sets no -O and no -march; x86-64 baseline has no FMA either way).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TURNER_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/lib" ]; then
    echo "build_ref.sh: $REF not present; keeping any prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/lib" "$TMP/src"
for f in "$REF"/lib/*; do ln -s "$f" "$TMP/lib/$(basename "$f")"; done
for f in "$REF"/src/*; do [ "$(basename "$f")" = geometry.h ] || ln -s "$f" "$TMP/src/$(basename "$f")"; done
for f in "$REF"/*.h "$REF"/*.cpp; do ln -s "$f" "$TMP/$(basename "$f")"; done
python3 - "$REF/src/geometry.h" "$TMP/src/geometry.h" <<'PY'
import sys
src = open(sys.argv[1]).read()
blocks = [
    ("    template <typename U> friend Point2<U> operator*(U s, const Point2<U>& p) {\n"
     "        return {s * p.x, s * p.y};\n    }\n",
     "using Point2i = Point2<int>;\n",
     "template <typename U> Point2<U> operator*(U s, const Point2<U>& p) {\n"
     "    return {s * p.x, s * p.y};\n}\n"),
    ("    template <typename U>\n    friend Normal3<U> operator*(U s, const Normal3<U>& n) {\n"
     "        return {s * n.x, s * n.y, s * n.z};\n    }\n",
     "using Normal3f = Normal3<float>;\n",
     "template <typename U> Normal3<U> operator*(U s, const Normal3<U>& n) {\n"
     "    return {s * n.x, s * n.y, s * n.z};\n}\n"),
]
for inclass, anchor, hoisted in blocks:
    if src.count(inclass) != 1 or src.count(anchor) != 1:
        sys.exit("build_ref.sh: geometry.h does not look as expected; refusing to patch")
    src = src.replace(inclass, "")
    src = src.replace(anchor, anchor + hoisted)
open(sys.argv[2], "w").write(src)
PY
CXX="${CXX:-g++}"
FLAGS="-std=c++14 -O2 -DNDEBUG -ffp-contract=off -fPIC -shared -pthread -w -s -nostdlib++ -I$HERE/ref_shims -I$TMP"
$CXX $FLAGS "$TMP/lib/kdtree.cpp" "$TMP/pathtracer.cpp" "$HERE/ref_driver.cpp" -l:libstdc++.so.6 -o "$OUT/libturner_ref_pathtracer.so"
$CXX $FLAGS "$TMP/lib/kdtree.cpp" "$TMP/raycaster.cpp" "$HERE/ref_driver.cpp" -l:libstdc++.so.6 -o "$OUT/libturner_ref_raycaster.so"
$CXX $FLAGS "$TMP/lib/kdtree.cpp" "$TMP/raytracer.cpp" "$HERE/ref_driver.cpp" -l:libstdc++.so.6 -o "$OUT/libturner_ref_raytracer.so"
# the as-shipped flags (CMakeLists.txt:3-7: no -O, asserts on), for comparability with README.md:22-36's 129 k rays/s
FLAGS_SHIPPED="-std=c++14 -fPIC -shared -pthread -w -s -nostdlib++ -I$HERE/ref_shims -I$TMP"
$CXX $FLAGS_SHIPPED "$TMP/lib/kdtree.cpp" "$TMP/pathtracer.cpp" "$HERE/ref_driver.cpp" -l:libstdc++.so.6 -o "$OUT/libturner_ref_pathtracer_shipped.so"
echo "built $OUT/libturner_ref_pathtracer.so $OUT/libturner_ref_raycaster.so"
