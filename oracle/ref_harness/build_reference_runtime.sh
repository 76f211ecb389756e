#!/bin/bash
# TEST INFRASTRUCTURE (optional, ~15 min on 8 cores): builds the reference's scalar_rgb runtime OUT OF TREE from the
# read-only sources in /root/reference with the reference's own CMake build and copies the binaries into oracle/_ref/
# (git-ignored; travels to the GPU box with gpurun). Nothing of the product depends on it:
#   * tests/golden/make_golden.py uses it (through oracle/ref_harness/replay_harness.cpp) to regenerate the fixtures,
#   * tests/test_mitsuba_plugin.py and `bench.py --impl reference` use it when present and skip / fall back otherwise.
# __graft_entry__.build() does NOT run this script (the reference needs cmake, Embree and a generated config.h, i.e. it
# does not compile from a few source files); it only rebuilds the small harness + plugin when oracle/_ref already exists.
# Workarounds: cmake 4 rejects old cmake_minimum_required in submodules; gcc 13 needs <cstdint>;
# /opt/gcc has no lto-wrapper while nanothread / drjit-core force IPO.
set -euo pipefail
REF=${REF:-/root/reference}
BUILD=${BUILD:-/tmp/refbuild}
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/../_ref"
OUT="$(cd "$HERE/../_ref" && pwd)"
mkdir -p "$BUILD" && cd "$BUILD"
cat > noipo.cmake <<'EOC'
set(CMAKE_C_COMPILE_OPTIONS_IPO "")
set(CMAKE_CXX_COMPILE_OPTIONS_IPO "")
set(CMAKE_C_LINK_OPTIONS_IPO "")
set(CMAKE_CXX_LINK_OPTIONS_IPO "")
EOC
cmake -G Ninja "$REF" -DCMAKE_BUILD_TYPE=Release -DMI_ENABLE_PYTHON=OFF -DMI_DEFAULT_VARIANTS="scalar_rgb" \
      -DCMAKE_POLICY_VERSION_MINIMUM=3.5 -DCMAKE_CXX_FLAGS="-include cstdint -include cstdio" \
      -DCMAKE_PROJECT_INCLUDE="$BUILD/noipo.cmake"
ninja -j"${JOBS:-6}"
cp -a mitsuba lib*.so plugins include "$OUT"/
# measured spectra of the named conductor materials (data, resolved as data/ior/<name>.{eta,k}.spd next to the executable)
mkdir -p "$OUT/data" && cp -a "$REF/resources/data/ior" "$OUT/data/"
make -C "$HERE"        # replay_harness, header_vectors, plugins/dopplertofpath_b200.so
echo "reference runtime + harness + plugin in $OUT"
