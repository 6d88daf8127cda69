#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (AMReX-Astro/Castro, read-only at
# /root/reference) into oracle/_ref/ so that the oracle restatement (oracle/ctu_oracle.cpp) and the
# CUDA path can be checked against the reference's own CPU executable, and so that
# `bench.py --impl reference` can time the reference's CPU implementation on the GPU box's cores.
#
# The reference hot path cannot be compiled "from a few source files": every hydro header pulls
# in headers that the reference *generates* at build time (state_indices.H, castro_params.H,
# prob_parameters.H, network_properties.H, extern_parameters.H, AMReX_Config.H).  The recipe
# therefore drives the reference's own GNU-make files from a scratch copy of the problem
# directory under /tmp (the reference tree itself is never written).  This is stated in DESIGN.md.
#
# Outputs (git-ignored, but shipped to the GPU box by gpurun):
#   oracle/_ref/Castro3d_sedov.ex        serial, -O3, no FMA  (parity oracle, golden generator)
#   oracle/_ref/Castro3d_sedov_omp.ex    OpenMP build         (CPU baseline for bench.py)
#   oracle/_ref/Castro3d_kh.ex           Kelvin-Helmholtz problem, serial (golden generator for the 2-passive path)
#   oracle/_ref/Castro3d_kh4.ex          (`kh4`) KH built with NETWORK_INPUTS=triple_alpha_plus_o.net (4 species + 1 advected scalar: NUM_STATE = 12,
#                                        BASELINE config[3]); golden generator for the 5-passive tile kernels
#   oracle/_ref/Castro3d_sod.ex          Sod shock-tube problem in 3-D, serial (golden generator for BASELINE config[0])
#   oracle/_ref/Castro3d_rt.ex           Rayleigh-Taylor problem with constant gravity, serial (golden generator for the source-term path).
#                                        The reference's RT initial-data header does not compile with DIM=3 (an unqualified `L_x`,
#                                        problem_initialize_state_data.H:50); the recipe qualifies that one identifier in the SCRATCH COPY of
#                                        the problem directory.  No file of the hot path (Source/hydro, Source/driver, Source/gravity) is touched.
#   oracle/_ref/Castro3d_noh.ex          (`noh`) Noh implosion problem, serial (golden generator for problem_bc_fill / inflow boundaries)
#   oracle/_ref/Castro2d_sedov.ex, Castro2d_kh.ex   (`sedov2d`, `kh2d`) DIM=2 builds, golden generators for the 2-D Cartesian path
#   oracle/_ref/fcompare.ex              AMReX plotfile comparator (DIM=3)
#   oracle/_ref/Castro3d_sedov_cuda.ex   (`gpu`)  the UNMODIFIED reference built with its own GPU back end (USE_CUDA=TRUE CUDA_ARCH=100):
#                                        Castro's stock unfused ParallelFor pipeline on the same B200 -- the "reference_gpu" baseline
#   oracle/_ref/Castro3d_sedov_stub.ex   (`stub`) the same GPU build with ONE source file swapped: integration/Castro_ctu_hydro.cpp (this
#                                        repository's binding of Castro::construct_ctu_hydro_source to libcastro_ctu.so) is copied into the
#                                        scratch problem directory, where Castro's VPATH finds it before Source/hydro/Castro_ctu_hydro.cpp.
#                                        Everything else (driver, FillPatch, clean_state, estdt, plotfile I/O) is the reference's own code.
#                                        Neither is built by `all` (20 min of nvcc); build() asks for them when they are missing.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
W=${W:-/tmp/castro_ref_build}
J=${J:-$(nproc)}
mkdir -p "$OUT" "$W"
[ -d "$REF" ] || { echo "no reference at $REF; keeping prebuilt oracle/_ref" ; exit 0; }

build() {  # $1 = problem dir name, $2 = output name, rest = make args
  local prob=$1 out=$2; shift 2
  local d="$W/${out}"
  if [ ! -d "$d" ]; then cp -r "$REF/Exec/hydro_tests/$prob" "$d"; fi
  ( cd "$d" && make CASTRO_HOME="$REF" USE_MPI=FALSE "$@" -j"$J" >build.log 2>&1 ) || { tail -30 "$d/build.log"; exit 1; }
  local exe; exe=$(ls -t "$d"/Castro*.ex | head -1)
  cp "$exe" "$OUT/$out.ex"
  case "$out" in *_cuda|*_stub) strip "$OUT/$out.ex" ;; esac      # symbol tables only (31 MB each): the gpurun snapshot is limited to 512 MiB
  echo "built $OUT/$out.ex"
}

what=${1:-all}
if [ "$what" = all ] || [ "$what" = serial ]; then build Sedov Castro3d_sedov DIM=3; fi
if [ "$what" = all ] || [ "$what" = omp ];    then build Sedov Castro3d_sedov_omp DIM=3 USE_OMP=TRUE; fi
# KH as shipped: NUMADV := 1 in its GNUmakefile -> NUM_STATE = 9 (one advected scalar + one species): pins the 2-passive path
if [ "$what" = all ] || [ "$what" = kh ];     then build KH Castro3d_kh DIM=3; fi
# Sod shock tube (BASELINE config[0]) built in 3-D: the tube along x, y or z exercises the three directional code paths
# KH with a 4-species network (BASELINE config[3]: NUMADV=1 + triple_alpha_plus_o.net -> NUM_STATE = 12): golden generator for the 5-passive path
if [ "$what" = kh4 ]; then build KH Castro3d_kh4 DIM=3 NETWORK_INPUTS=triple_alpha_plus_o.net; fi
# Sod_stellar: the shock tube with the Helmholtz (stellar) EOS and the 3-species ignition network; golden generator
if [ "$what" = sodstellar ]; then build Sod_stellar Castro3d_sodstellar DIM=3 USE_MHD=FALSE; fi
# ... and in the dimensionalities the problem ships with (GNUmakefile: DIM = 2; the README runs it in 1-D): golden generators for the Helmholtz EOS in 1-D / 2-D
if [ "$what" = sodstellar2d ]; then build Sod_stellar Castro2d_sodstellar DIM=2 USE_MHD=FALSE; fi
if [ "$what" = sodstellar1d ]; then build Sod_stellar Castro1d_sodstellar DIM=1 USE_MHD=FALSE; fi
# ... and with the gravity module compiled in (USE_GRAV=TRUE, gravity.gravity_type = ConstantGrav): golden generator for the Helmholtz EOS
# together with the constant-gravity source terms and the source predictor (the literal combination of BASELINE config[4])
if [ "$what" = sodstellar_grav ]; then build Sod_stellar Castro3d_sodstellar_grav DIM=3 USE_MHD=FALSE USE_GRAV=TRUE; fi
if [ "$what" = sodstellar2d_grav ]; then build Sod_stellar Castro2d_sodstellar_grav DIM=2 USE_MHD=FALSE USE_GRAV=TRUE; fi
# the literal BASELINE config[0]: Sod in ONE dimension (DIM=1), golden generator for the 1-D path
if [ "$what" = sod1d ]; then build Sod Castro1d_sod DIM=1; fi
# 2-D Cartesian path (AMREX_SPACEDIM == 2): Sedov (cylindrical blast in Cartesian coordinates, inputs.2d.cyl_in_cartcoords) and KH (inputs.2d)
if [ "$what" = sedov1d ]; then build Sedov Castro1d_sedov DIM=1; fi      # 1-D spherical / cylindrical Sedov (inputs.1d.sph, inputs.1d.cyl): the non-Cartesian 1-D terms
if [ "$what" = sedov2d ]; then build Sedov Castro2d_sedov DIM=2; fi
if [ "$what" = kh2d ]; then build KH Castro2d_kh DIM=2; fi
if [ "$what" = rt2d ]; then build RT Castro2d_rt DIM=2; fi      # Rayleigh-Taylor as shipped (inputs_2d): constant gravity along y
if [ "$what" = all ] || [ "$what" = sod ];    then build Sod Castro3d_sod DIM=3; fi
# Noh implosion: the one hydro test whose outer ghost zones come from problem_bc_fill (time-dependent inflow); golden generator for the bc hook
if [ "$what" = noh ]; then build Noh Castro3d_noh DIM=3; fi
if [ "$what" = all ] || [ "$what" = rt ]; then
  if [ ! -d "$W/Castro3d_rt" ]; then cp -r "$REF/Exec/hydro_tests/RT" "$W/Castro3d_rt"; fi
  sed -i 's/0.5e0_rt \* L_x/0.5e0_rt * problem::L_x/' "$W/Castro3d_rt/problem_initialize_state_data.H"
  build RT Castro3d_rt DIM=3
fi
if [ "$what" = gpu ]; then build Sedov Castro3d_sedov_cuda DIM=3 USE_CUDA=TRUE CUDA_ARCH=100; fi
if [ "$what" = stub ]; then
  REPO="$(cd "$HERE/.." && pwd)"
  d="$W/Castro3d_sedov_stub"
  if [ ! -d "$d" ]; then
    if [ -d "$W/Castro3d_sedov_cuda" ]; then cp -r "$W/Castro3d_sedov_cuda" "$d"; rm -f "$d"/Castro*.ex;      # reuse the objects of the stock GPU build
    else cp -r "$REF/Exec/hydro_tests/Sedov" "$d"; fi
  fi
  cp "$REPO/integration/Castro_ctu_hydro.cpp" "$d/Castro_ctu_hydro.cpp"
  if ! grep -q castro_ctu "$d/GNUmakefile"; then
    sed -i "s|^include \$(CASTRO_HOME)/Exec/Make.Castro|INCLUDE_LOCATIONS += $REPO/include\nLIBRARIES += -L$REPO/castro_b200 -lcastro_ctu -Xlinker -rpath -Xlinker /root/repo/castro_b200\n&|" "$d/GNUmakefile"
  fi
  [ -f "$REPO/castro_b200/libcastro_ctu.so" ] || make -C "$REPO/castro_b200/csrc" -j"$J"
  build Sedov Castro3d_sedov_stub DIM=3 USE_CUDA=TRUE CUDA_ARCH=100
fi
if [ "$what" = all ] || [ "$what" = fcompare ]; then
  d="$W/fcompare"; mkdir -p "$d"; cp -r "$REF/external/amrex/Tools/Plotfile/"* "$d"/
  ( cd "$d" && make AMREX_HOME="$REF/external/amrex" programs=fcompare DIM=3 USE_MPI=FALSE -j"$J" >build.log 2>&1 ) || { tail -30 "$d/build.log"; exit 1; }
  cp "$d"/fcompare.gnu.ex "$OUT/fcompare.ex"; echo "built $OUT/fcompare.ex"
fi
