#!/bin/bash
# build_ref_pic.sh -- recipe for oracle/_ref/libref_pic.so: the reference's OWN PIC core as a one-rank checker of the oracle.
# OURS (test infrastructure).  Usage: build_ref_pic.sh <reference tree> <output .so>
#
#  1. a scratch copy of the reference's sources under $TMPDIR (its own configuration step rewrites the tree it runs in;
#     /root/reference is read-only and no reference source enters this repository),
#  2. the reference's own Perl configuration for its ECSIM test `input/test/fast-wave.input` (ampsConfig.pl -no-compile): it
#     generates build/ with _BLOCK_CELLS_ 16,8,4, Lapenta2017 as the mover, ECSIM as the field solver, the periodic mode,
#  3. every translation unit the reference's makefiles list for that configuration (src/pic, general, meshAMR, interface,
#     species, models/*; the SWMF fluid coupler pic_fluid.cpp / pic_swmf.cpp excepted) compiled with g++ -O3 -ffp-contract=off
#     against our stand-ins for mpi.h (one rank) and the three un-vendored SWMF `share` headers (oracle/ref_pic/, oracle/ref_mesh/mpi.h),
#  4. linked with oracle/ref_pic/ref_pic_shim.cpp (includes the reference's test/srcFastWave/main.cpp for the initial
#     conditions) and oracle/ref_pic/mpi_single.cpp into the shared library.
set -e
REF=${1:-/root/reference}
OUT=${2:-$(dirname "$0")/../_ref/libref_pic.so}
HERE=$(cd "$(dirname "$0")" && pwd)
S=${TMPDIR:-/tmp}/amps_b200_ref_pic_$$
JOBS=${JOBS:-$(nproc)}
mkdir -p "$S" "$(dirname "$OUT")"
OUT=$(cd "$(dirname "$OUT")" && pwd)/$(basename "$OUT")
trap 'rm -rf "$S"' EXIT
cp -r "$REF"/src "$REF"/srcInterface "$REF"/utility "$REF"/input "$REF"/test "$REF"/*.pl "$REF"/*.pm "$REF"/Makefile* "$S"/ 2>/dev/null || true
cd "$S"
cp input/test/fast-wave.* input/species.input .
touch .amps.conf .general.conf
perl ampsConfig.pl -input fast-wave.input -no-compile > config.log 2>&1 || { tail -20 config.log; exit 1; }
# REF_PIC_VARIANT=gk: the same configuration with the gyrokinetic model switched on in the GENERATED picGlobal.dfn of the scratch tree
# (what `ampsConfigLib::RedefineMacro` does for an input file that asks for it): particles carry the magnetic moment, v_parallel and
# v_normal, ProcessCell takes its guiding-centre branch for the species of UseGuidingCenterSpeciesTable, MoveParticles calls
# the shim's ref_pic_user_mover, which is PIC::GYROKINETIC::Mover (guiding-centre mover for those species, Lapenta2017 for the rest) or,
# on request, GuidingCenter::Mover_SecondOrder for the guiding-centre species; ProcessCell also samples the species
# moments on the corners (_PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_), which CorrectParticleLocation of the div-E correction reads
if [ "${REF_PIC_VARIANT:-}" = "gk" ]; then
  sed -i -E 's/^#define _USE_MAGNETIC_MOMENT_ .*/#define _USE_MAGNETIC_MOMENT_ _PIC_MODE_ON_/; s/^#define _PIC_GYROKINETIC_MODEL_MODE_ .*/#define _PIC_GYROKINETIC_MODEL_MODE_ _PIC_MODE_ON_/; s/^#define _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ .*/#define _PIC_FIELD_SOLVER_SAMPLE_SPECIES_ON_CORNER_ _PIC_MODE_ON_/; s/^#define _PIC_PARTICLE_MOVER__MOVE_PARTICLE_TIME_STEP_\(ptr,LocalTimeStep,node\).*/#define _PIC_PARTICLE_MOVER__MOVE_PARTICLE_TIME_STEP_(ptr,LocalTimeStep,node) ref_pic_user_mover(ptr,LocalTimeStep,(void*)(node));/' build/pic/picGlobal.dfn
  grep -n "_USE_MAGNETIC_MOMENT_ \|_PIC_GYROKINETIC_MODEL_MODE_ \|MOVE_PARTICLE_TIME_STEP_(" build/pic/picGlobal.dfn
fi
# the object lists of the reference's own makefiles
python3 - "$S" <<'PY' > tus.txt
import os, re, sys
S = sys.argv[1]
for d in ['pic', 'general', 'meshAMR', 'interface', 'species', 'models/exosphere', 'models/dust', 'models/surface', 'models/sputtering',
          'models/charge_exchange', 'models/electron_impact', 'models/photolytic_reactions']:
    mk = os.path.join(S, 'build', d, 'makefile')
    if not os.path.exists(mk):
        continue
    for o in sorted(set(re.findall(r'([A-Za-z0-9_/\-]+)\.o\b', open(mk).read().replace('\\\n', ' ')))):
        p = os.path.join(d, o + '.cpp')
        if os.path.exists(os.path.join(S, 'build', p)) and os.path.basename(p) not in ('pic_fluid.cpp', 'pic_swmf.cpp'):
            print(p)
PY
B="$S/build"
INC="-include $HERE/ref_pic_stubs.h -I$HERE -I$S/test/srcFastWave -I$B/pic -I$B/general -I$B/meshAMR -I$B/interface -I$B/models/exosphere -I$B/models/dust -I$B/species -I$B/models/surface -I$B/models/sputtering -I$B/models/charge_exchange -I$B/models/electron_impact -I$B/models/photolytic_reactions -I$S/srcInterface -I$HERE/../ref_mesh"
# -O3 without FMA contraction: the same library is the checker of the oracle (bit-identical results, tests/test_reference_ecsim.py) and
# the reference-code CPU arm of bench.py
FLAGS="-std=c++17 -w ${REF_PIC_OPT:--O3 -ffp-contract=off} -fPIC"
mkdir -p obj
cat > cc.sh <<EOS
#!/bin/bash
f=\$1; o=$S/obj/\$(echo \$f | tr '/' '_' | sed 's/\.cpp\$/.o/')
cd $B/\$(dirname \$f) && g++ $FLAGS $INC -c \$(basename \$f) -o \$o > \$o.log 2>&1 || { echo "FAILED \$f"; grep -m3 error \$o.log; exit 1; }
EOS
chmod +x cc.sh
xargs -P "$JOBS" -n 1 ./cc.sh < tus.txt
(cd "$B/pic" && g++ $FLAGS -fno-access-control $INC -c "$HERE/ref_pic_shim.cpp" -o "$S/obj/ref_pic_shim.o")
g++ $FLAGS -I"$HERE" -c "$HERE/gmres_single.cpp" -o "$S/obj/gmres_single.o"
g++ $FLAGS -I"$HERE" -I"$HERE/../ref_mesh" -c "$HERE/mpi_single.cpp" -o "$S/obj/mpi_single.o"
g++ -shared -o "$OUT" obj/*.o -lpthread
echo "built $OUT from $(wc -l < tus.txt) reference translation units"
