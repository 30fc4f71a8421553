#!/bin/sh
# Builds oracle/_ref/libref_pairs.so from the REFERENCE's own scalar code (MATLAB-codegen bodies of the
# pair-distance gradients/Hessians, the EE cross-norm derivatives, the mollifier polynomial and the C2 barrier)
# compiled from the sources where they lie under /root/reference.  The line ranges below are the pure-`double`
# overloads (no Eigen inside); they are streamed through sed into a scratch file OUTSIDE the repo, compiled, and
# the scratch file is deleted -- no reference source is ever copied into this repository.  Only the .so lands in
# oracle/_ref/ (git-ignored, travels to the GPU box).  The rest of the reference (Eigen/TBB/libigl/CCD-Wrapper
# dependent) is unbuildable here: see DESIGN.md.
set -e
REF=/root/reference/src
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
M="$REF/CollisionObject/MeshCollisionUtils.hpp"
B="$REF/Utils/BarrierFunctions.hpp"
{
  echo '#include <cmath>'
  echo 'namespace refsrc {'
  sed -n '235,287p;300,620p;696,757p;772,1217p;1298,1377p;1392,2002p;2419,2477p;2494,2762p;2834,2849p' "$M"
  sed -n '56,83p' "$B"
  echo '}'
  cat <<'SHIM'
extern "C" {
#define V9 v[0],v[1],v[2],v[3],v[4],v[5],v[6],v[7],v[8]
#define V12 V9,v[9],v[10],v[11]
void ref_g_PE(const double* v, double* g) { refsrc::g_PE(V9, g); }
void ref_H_PE(const double* v, double* H) { refsrc::H_PE(V9, H); }
void ref_g_PT(const double* v, double* g) { refsrc::g_PT(V12, g); }
void ref_H_PT(const double* v, double* H) { refsrc::H_PT(V12, H); }
void ref_g_EE(const double* v, double* g) { refsrc::g_EE(V12, g); }
void ref_H_EE(const double* v, double* H) { refsrc::H_EE(V12, H); }
void ref_EEcross_g(const double* v, double* g) { refsrc::computeEECrossSqNormGradient(V12, g); }
void ref_EEcross_H(const double* v, double* H) { refsrc::computeEECrossSqNormHessian(V12, H); }
void ref_q(double x, double eps, double* q, double* qg, double* qH) { refsrc::compute_q(x, eps, *q); refsrc::compute_q_g(x, eps, *qg); refsrc::compute_q_H(x, eps, *qH); }
void ref_barrier(double d, double dHat, double* b, double* g, double* H) { refsrc::b_C2(d, dHat, *b); refsrc::g_bC2(d, dHat, *g); refsrc::H_bC2(d, dHat, *H); }
}
SHIM
} > "$TMP/ref_pairs.cpp"
$CXX -O2 -std=c++17 -fPIC -ffp-contract=off -shared -o "$OUT/libref_pairs.so" "$TMP/ref_pairs.cpp"
echo "built $OUT/libref_pairs.so"
