#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
#
# Compiles the UNMODIFIED GRANIITTI reference sources where they lie under /root/reference into
# oracle/_ref/ (git-ignored, but NOT gpurun-ignored: the binaries and their run tree travel to the GPU
# box, where /root/reference does not exist). Nothing is copied into the tracked tree.
#
# Produces
#   oracle/_ref/gra/bin/gr        stock reference CLI (CPU baseline for `bench.py --impl reference`)
#   oracle/_ref/gra/bin/ref_eval  per-point harness (oracle/ref_harness/ref_eval.cc) that calls the
#                                 reference's MProcess::EventWeight on caller supplied points with the
#                                 internal uniforms (MRandom::U) read from a tape
#   oracle/_ref/gra/{modeldata,input,tests,MG5cards,...}   run tree: DATA files the binaries read at run
#                                 time (JSON tunes, PDG table, steering cards), copied so they travel
#
# The reference's own Makefile is not used (it needs `bc`, HepMC3 and LHAPDF installs). HepMC3 is
# replaced by the header stub in oracle/ref_harness/hepmc3_stub; LHAPDF 6.2.3 is built from the tarball
# shipped inside the reference (install/LHAPDF-6.2.3.tar.gz). Flags follow reference Makefile:195-205.
set -euo pipefail
R=${GRB_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$R/src" ]; then echo "build_ref: $R not present, nothing to do"; exit 0; fi
mkdir -p "$OUT/obj" "$OUT/lhapdf_obj" "$OUT/gra/bin"

# ---- 1. LHAPDF static library ------------------------------------------------------------------------
L=$OUT/LHAPDF-6.2.3
if [ ! -f "$OUT/libLHAPDF.a" ]; then
  [ -d "$L" ] || tar xf "$R/install/LHAPDF-6.2.3.tar.gz" -C "$OUT"
  # per-file compile helper (keeps the string-valued -D macros out of nested shell quoting)
  cat > "$OUT/lhapdf_obj/cc1.sh" <<EOS
#!/bin/sh
f=\$1; b=\$(basename \$f); b=\${b%.*}; extra=""
[ "\$b" = Info ] && extra="-DYAML=LHAPDF_YAML -DYAML_NAMESPACE=LHAPDF_YAML"
exec g++ -std=c++11 -O2 -fPIC -w -I$L/include -I$L/src/yamlcpp -I$L/src/yamlcpp/yaml-cpp -I$L/src \\
  '-DLHAPDF_DATA_PREFIX="$OUT/share"' '-DLHAPDF_INSTALL_PREFIX="$OUT"' \$extra -c \$f -o $OUT/lhapdf_obj/\$b.o
EOS
  chmod +x "$OUT/lhapdf_obj/cc1.sh"
  ls $L/src/*.cc $L/src/yamlcpp/*.cpp | xargs -P "$JOBS" -n 1 "$OUT/lhapdf_obj/cc1.sh"
  ar rcs "$OUT/libLHAPDF.a" "$OUT"/lhapdf_obj/*.o
fi

# ---- 2. reference objects (unmodified sources, Makefile:195-205 flags) ---------------------------------
CXXF="-std=c++17 -O3 -DNDEBUG -ftree-vectorize -fno-signed-zeros -fPIC -w \
  -I$HERE/ref_harness/hepmc3_stub -I$L/include -I$R -I$R/include -I$R/libs -I$R/libs/FTensor -I$R/libs/Eigen/unsupported"
( cd "$OUT/obj"
  ls $R/src/*.cc $R/src/Amplitude/*.cc $R/src/Program/gr.cc | xargs -P "$JOBS" -I{} sh -c \
    'f={}; b=$(basename $f .cc); [ -f $b.o ] && [ $b.o -nt $f ] || g++ '"$CXXF"' -c $f -o $b.o' )

# ---- 3. stock CLI -------------------------------------------------------------------------------------
g++ "$OUT"/obj/*.o -o "$OUT/gra/bin/gr" -L"$OUT" -lLHAPDF -pthread -lrt

# ---- 4. per-point harness: same objects, but MRandom::U made weak so the tape version overrides it -----
mkdir -p "$OUT/obj_h"
cp "$OUT/obj/MRandom.o" "$OUT/obj_h/MRandom_weakU.o"
objcopy --weaken-symbol=_ZN3gra7MRandom1UEdd "$OUT/obj_h/MRandom_weakU.o"
g++ $CXXF -c "$HERE/ref_harness/ref_eval.cc" -o "$OUT/obj_h/ref_eval.o"
OBJS=$(ls "$OUT"/obj/*.o | grep -v -e '/gr.o$' -e '/MRandom.o$')
g++ "$OUT/obj_h/ref_eval.o" $OBJS "$OUT/obj_h/MRandom_weakU.o" -o "$OUT/gra/bin/ref_eval" -L"$OUT" -lLHAPDF -pthread -lrt

# ---- 4a. the reference's header-only VEGAS grid refinement (MVEGAS.h), for the direct OptimizeGrid goldens -----------
g++ -std=c++17 -O2 -I$R/include "$HERE/ref_harness/ref_vegas.cc" -o "$OUT/gra/bin/ref_vegas"

# ---- 4b. INTEGRATION.md route A, compiled: the reference's steering + the product's C ABI (needs graniitti_b200/libgrb.so) --
LIBGRB="$HERE/../graniitti_b200/libgrb.so"
if [ -f "$LIBGRB" ]; then
  g++ $CXXF -I"$HERE/../include" -c "$HERE/ref_harness/gr_b200.cc" -o "$OUT/obj_h/gr_b200.o"
  g++ "$OUT/obj_h/gr_b200.o" $OBJS "$OUT/obj/MRandom.o" -o "$OUT/gra/bin/gr_b200" -L"$OUT" -lLHAPDF -pthread -lrt \
      "$LIBGRB" -Wl,-rpath,'$ORIGIN/../../../../graniitti_b200'
fi

# ---- 5. run tree (the program resolves data relative to <binary>/../ ; reference tree is read-only) ----
cd "$OUT/gra"
mkdir -p eikonal output sudakov figs
for d in modeldata MG5cards input fitcard; do rm -rf "$d"; cp -r "$R/$d" "$d"; done
mkdir -p tests
for d in processes run_atlas_multi run_cms_mumu_low_high run_rhorho_phiphi LHC_TEVATRON_RHIC; do
  rm -rf "tests/$d"; [ -d "$R/tests/$d" ] && cp -r "$R/tests/$d" "tests/$d"
done
# a second tune for the goldens of PARAM_STRUCTURE.QED_alpha != "ZERO" (MG5 param-card coupling): TUNE0 with that one key changed
rm -rf modeldata/TUNE0_MG; cp -r modeldata/TUNE0 modeldata/TUNE0_MG
sed -i 's/"QED_alpha"[ ]*:[ ]*"ZERO"/"QED_alpha" : "MG"/' modeldata/TUNE0_MG/GENERAL.json
echo "build_ref: OK -> $OUT/gra/bin/{gr,ref_eval}"
