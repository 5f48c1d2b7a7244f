// rb_bsdf.cuh -- Klems-matrix BSDF data on the device: the queries m_bsdf() makes of the reference's BSDF library,
// on the tables rb_bsdf.cpp lays out (grayscale matrices, every cumulative distribution precomputed).
//
// Restated reference functions:
//   fo_getvec / fo_getndx / io_getohm and the fi / bi / bo variants   src/radiance/common/bsdf_m.c:138-291
//   mBSDF_color (grayscale)     src/radiance/common/bsdf_m.c:293-314
//   SDgetMtxBSDF                src/radiance/common/bsdf_m.c:728-751
//   SDqueryMtxProjSA            src/radiance/common/bsdf_m.c:753-800
//   SDgetMtxCDist / SDsampMtxCDist  src/radiance/common/bsdf_m.c:833-916 (tables instead of a cache list)
//   SDsizeBSDF                  src/radiance/common/bsdf.c:585-649
//   SDevalBSDF                  src/radiance/common/bsdf.c:651-696
//   SDdirectHemi                src/radiance/common/bsdf.c:698-744
//   SDsampComponent             src/radiance/common/bsdf.c:493-536
//   SDcompXform / SDinvXform / SDmapDir   src/radiance/common/bsdf.c:836-896
// Not restated: SDmultiSamp()'s Hilbert-curve split of the left-over random variable into the two coordinates of a
// patch (bsdf.c:540-569) -- two independent uniforms are drawn, the same distribution.
#pragma once
#include "rb_device.cuh"

namespace rb {

enum : int { SDQ_MIN = 1, SDQ_MAX = 2 };

// fvect.c:16-23
__device__ __forceinline__ double bsdf_Acos(double x) {
    if (x <= -1. + RB_FTINY * RB_FTINY) return RB_PI;
    if (x >= 1. - RB_FTINY * RB_FTINY) return 0.;
    return acos(x);
}

// fo_getndx(): index of the front-exiting vector (vx, vy, vz), -1 outside the basis
__device__ __noinline__ int kb_getndx(const BsdfBasis& ab, double vx, double vy, double vz) {
    if ((vz < 0) | (vz > 1.00001)) return -1;
    const double pol = 180.0 / RB_PI * bsdf_Acos(vz);
    double azi = 180.0 / RB_PI * atan2(vy, vx);
    if (azi < 0.0) azi += 360.0;
    int li;
    for (li = 1; ab.tmin[li] <= pol; li++)
        if (!ab.nphis[li]) return -1;
    --li;
    int ndx = (int)((1. / 360.) * azi * ab.nphis[li] + 0.5);
    if (ndx >= ab.nphis[li]) ndx = 0;
    while (li--) ndx += ab.nphis[li];
    return ndx;
}

// io_getohm(): projected solid angle of patch ndx, -1 for a bad index
__device__ __noinline__ double kb_ohm(const BsdfBasis& ab, int ndx) {
    if ((ndx < 0) | (ndx >= ab.nangles)) return -1.;
    int li;
    for (li = 0; ndx >= ab.nphis[li]; li++) ndx -= ab.nphis[li];
    const double c0 = cos(RB_PI / 180. * ab.tmin[li]), c1 = cos(RB_PI / 180. * ab.tmin[li + 1]);
    return RB_PI * (c0 * c0 - c1 * c1) / (double)ab.nphis[li];
}

// fo_getvec(): a direction inside patch ndx (rx0, rx1 uniform in [0,1))
__device__ __noinline__ void kb_getvec(const BsdfBasis& ab, int ndx, double rx0, double rx1, double v[3]) {
    int li;
    for (li = 0; ndx >= ab.nphis[li]; li++) ndx -= ab.nphis[li];
    const double c0 = cos(RB_PI / 180. * ab.tmin[li]), c1 = cos(RB_PI / 180. * ab.tmin[li + 1]);
    double d = (1. - rx0) * (c0 * c0) + rx0 * (c1 * c1);
    v[2] = d = sqrt(d);
    const double azi = 2. * RB_PI * (ndx + rx1 - .5) / ab.nphis[li];
    d = sqrt(1. - d * d);
    v[0] = cos(azi) * d;
    v[1] = sin(azi) * d;
}

// The four matrices of a file differ in which hemispheres their incident and exiting bases cover
// (load_bsdf_data(), bsdf_m.c:478-498): tf = fi/bo, tb = bi/fo, rf = fi/fo, rb = bi/bo.
__device__ __forceinline__ bool comp_in_front(int k) { return k == BC_TF || k == BC_RF; }     // incident basis: fi, else bi
__device__ __forceinline__ bool comp_out_front(int k) { return k == BC_TB || k == BC_RF; }    // exiting basis: fo, else bo

struct BsdfRef {
    const BsdfRec* rec;
    const BsdfBasis* bases;
    const unsigned* pool;
};

__device__ __forceinline__ int comp_incndx(const BsdfRef& B, int k, const double v[3]) {
    const BsdfBasis& ab = B.bases[B.rec->c[k].ib];
    return comp_in_front(k) ? kb_getndx(ab, -v[0], -v[1], v[2]) : kb_getndx(ab, -v[0], -v[1], -v[2]);
}
__device__ __forceinline__ int comp_outndx(const BsdfRef& B, int k, const double v[3]) {
    const BsdfBasis& ab = B.bases[B.rec->c[k].ob];
    return comp_out_front(k) ? kb_getndx(ab, v[0], v[1], v[2]) : kb_getndx(ab, v[0], v[1], -v[2]);
}
__device__ __forceinline__ double comp_incohm(const BsdfRef& B, int k, int i) { return kb_ohm(B.bases[B.rec->c[k].ib], i); }
__device__ __forceinline__ double comp_outohm(const BsdfRef& B, int k, int o) { return kb_ohm(B.bases[B.rec->c[k].ob], o); }

// mBSDF_color(), grayscale
__device__ __forceinline__ float comp_color(const BsdfRef& B, int k, int i, int o) {
    const BsdfComp& c = B.rec->c[k];
    float coef = __uint_as_float(__ldg(&B.pool[c.mtx + (unsigned)(o * c.ninc + i)]));
    double d = 2 * c.ninc / (i + .22545) + 4 * c.nout / (o + .70281);
    d -= (int)d;
    coef = (float)(coef * (1. + 6e-4 * (d - .5)));
    return coef;
}

// SDgetMtxBSDF(): false = nothing from this component
__device__ __noinline__ bool comp_get(const BsdfRef& B, int k, const double inVec[3], const double outVec[3], float& coef) {
    int i = comp_incndx(B, k, inVec), o = comp_outndx(B, k, outVec);
    if ((i < 0) & (o < 0)) {                     // try reciprocity
        i = comp_incndx(B, k, outVec);
        o = comp_outndx(B, k, inVec);
    }
    if ((i < 0) | (o < 0)) return false;
    coef = comp_color(B, k, i, o);
    return true;
}

// SDqueryMtxProjSA(); v2 == nullptr asks about v1 alone
__device__ __noinline__ void comp_query(const BsdfRef& B, int k, double psa[2], const double v1[3], const double* v2, int qflags) {
    const bool same = v2 == nullptr;
    if (same) v2 = v1;
    double out_psa = comp_outohm(B, k, comp_outndx(B, k, v1));
    double inc_psa = comp_incohm(B, k, comp_incndx(B, k, v2));
    if (!same & (out_psa <= 0) & (inc_psa <= 0)) {
        inc_psa = comp_outohm(B, k, comp_outndx(B, k, v2));
        out_psa = comp_incohm(B, k, comp_incndx(B, k, v1));
    }
    if (qflags == SDQ_MIN + SDQ_MAX) {
        if (inc_psa > psa[1]) psa[1] = inc_psa;
        if (out_psa > psa[1]) psa[1] = out_psa;
    }
    if ((inc_psa > 0) & (inc_psa < psa[0])) psa[0] = inc_psa;     // SDqueryMin part
    if ((out_psa > 0) & (out_psa < psa[0])) psa[0] = out_psa;
}

// the transmission matrix a ray arriving on the front (or back) side uses: tf if present else tb, and vice versa
__device__ __forceinline__ int bsdf_tcomp(const BsdfRec& R, bool front) {
    if (front) return R.c[BC_TF].present ? BC_TF : R.c[BC_TB].present ? BC_TB : -1;
    return R.c[BC_TB].present ? BC_TB : R.c[BC_TF].present ? BC_TF : -1;
}
__device__ __forceinline__ int bsdf_rcomp(const BsdfRec& R, bool front) {
    const int k = front ? BC_RF : BC_RB;
    return R.c[k].present ? k : -1;
}

// SDsizeBSDF() with qflags = SDqueryMin (projSA[0]) or SDqueryMin + SDqueryMax (projSA[0], projSA[1])
__device__ __noinline__ void sd_size(const BsdfRef& B, double projSA[2], const double v1[3], const double* v2, int qflags) {
    if (qflags == SDQ_MIN + SDQ_MAX) projSA[1] = .0;
    projSA[0] = 10.;
    const bool front = v1[2] > 0;
    int rk = bsdf_rcomp(*B.rec, front), tk = bsdf_tcomp(*B.rec, front);
    if (v2 != nullptr) {                         // bidirectional?
        if ((v1[2] > 0) ^ (v2[2] > 0)) rk = -1; else tk = -1;
    }
    if (rk >= 0) comp_query(B, rk, projSA, v1, v2, qflags);
    if (tk >= 0) comp_query(B, tk, projSA, v1, v2, qflags);
    if ((rk < 0) & (tk < 0)) {                   // all diffuse?
        projSA[0] = RB_PI;
        if (qflags == SDQ_MIN + SDQ_MAX) projSA[1] = RB_PI;
    } else if (qflags == SDQ_MIN + SDQ_MAX && projSA[0] > projSA[1])
        projSA[0] = projSA[1];
}

// SDevalBSDF(), cieY only (grayscale data: the colour is the equal-energy white)
__device__ __noinline__ double sd_eval(const BsdfRef& B, const double inVec[3], const double outVec[3]) {
    const bool inFront = inVec[2] > 0, outFront = outVec[2] > 0;
    const BsdfRec& R = *B.rec;
    double y; int k;
    if (inFront & outFront) { y = R.lamb[0]; k = bsdf_rcomp(R, true); }
    else if (!(inFront | outFront)) { y = R.lamb[1]; k = bsdf_rcomp(R, false); }
    else if (inFront) { y = R.lamb[2]; k = bsdf_tcomp(R, true); }
    else { y = R.lamb[3]; k = bsdf_tcomp(R, false); }
    y *= 1. / RB_PI;
    float coef;
    if (k >= 0 && comp_get(B, k, inVec, outVec, coef)) y += coef;
    return y;
}

// SDgetMtxCDist(): which cumulative table serves inVec.  Returns false when the vector is in neither basis.
__device__ __forceinline__ bool comp_cdist(const BsdfRef& B, int k, const double inVec[3], int& indx, bool& reverse) {
    indx = comp_incndx(B, k, inVec);
    reverse = false;
    if (indx >= 0) return true;
    indx = comp_outndx(B, k, inVec);             // try reciprocity
    reverse = true;
    return indx >= 0;
}
__device__ __forceinline__ double comp_ctotal(const BsdfRef& B, int k, int indx, bool reverse) {
    const BsdfComp& c = B.rec->c[k];
    return __ldg((const double*)(B.pool + (reverse ? c.rctot : c.ctot)) + indx);
}

// SDdirectHemi() for sflags = SDsampSpR (xmit = false) or SDsampSpT (xmit = true): non-diffuse part only
__device__ __noinline__ double sd_direct_hemi(const BsdfRef& B, const double inVec[3], bool xmit) {
    const bool front = inVec[2] > 0;
    const int k = xmit ? bsdf_tcomp(*B.rec, front) : bsdf_rcomp(*B.rec, front);
    if (k < 0) return 0.;
    int indx; bool rev;
    if (!comp_cdist(B, k, inVec, indx, rev)) return 0.;
    return comp_ctotal(B, k, indx, rev);
}

// SDsampComponent(): ioVec in = incident, out = sampled direction; returns the sample's cieY (0: nothing to sample)
__device__ __noinline__ double comp_sample(const BsdfRef& B, int k, double ioVec[3], double randX, double rx0, double rx1) {
    int indx; bool rev;
    double cieY = 0;
    if (comp_cdist(B, k, ioVec, indx, rev)) cieY = comp_ctotal(B, k, indx, rev);
    if (cieY <= 1e-6) { ioVec[0] = ioVec[1] = ioVec[2] = 0.; return 0.; }
    const BsdfComp& c = B.rec->c[k];
    const int calen = rev ? c.ninc : c.nout;
    const unsigned* carr = B.pool + (rev ? c.rcdf : c.cdf) + (size_t)indx * (calen + 1);
    const double maxval = 4294967295.0;
    const unsigned target = (unsigned)(randX * maxval);
    int i, ilower = 0, iupper = calen;           // SDsampMtxCDist(): binary search
    while ((i = (iupper + ilower) >> 1) != ilower)
        if (target >= __ldg(&carr[i])) ilower = i; else iupper = i;
    // (the reference turns the position inside the table step into the patch's two coordinates; rx0, rx1 stand for them)
    const BsdfBasis& ab = B.bases[rev ? c.ib : c.ob];
    kb_getvec(ab, i, rx0, rx1, ioVec);
    const bool front = rev ? comp_in_front(k) : comp_out_front(k);
    if (rev) {                                   // ib_vec: fi (-x, -y, z) or bi (-x, -y, -z)
        ioVec[0] = -ioVec[0]; ioVec[1] = -ioVec[1];
        if (!front) ioVec[2] = -ioVec[2];
    } else if (!front)                           // ob_vec: fo or bo (x, y, -z)
        ioVec[2] = -ioVec[2];
    return cieY;
}

// SDcompXform(): world -> BSDF coordinates from the surface normal and the up vector; false = illegal orientation
__device__ __forceinline__ double bsdf_normalize(double v[3]) {      // fvect.c:130-157, as normalize3() of rb_shade.cuh
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (d == 0.0) return 0.0;
    double len;
    if ((d <= 1.0 + RB_FTINY) & (d >= 1.0 - RB_FTINY)) { len = 0.5 + 0.5 * d; d = 2.0 - len; }
    else { len = sqrt(d); d = 1.0 / len; }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}
__device__ __forceinline__ bool sd_comp_xform(double m[3][3], const double sNrm[3], const double uVec[3]) {
    for (int k = 0; k < 3; k++) m[2][k] = sNrm[k];
    if (bsdf_normalize(m[2]) == 0) return false;
    m[0][0] = uVec[1] * m[2][2] - uVec[2] * m[2][1];
    m[0][1] = uVec[2] * m[2][0] - uVec[0] * m[2][2];
    m[0][2] = uVec[0] * m[2][1] - uVec[1] * m[2][0];
    if (bsdf_normalize(m[0]) == 0) return false;
    m[1][0] = m[2][1] * m[0][2] - m[2][2] * m[0][1];
    m[1][1] = m[2][2] * m[0][0] - m[2][0] * m[0][2];
    m[1][2] = m[2][0] * m[0][1] - m[2][1] * m[0][0];
    return true;
}
__device__ __forceinline__ bool sd_inv_xform(double im[3][3], const double v[3][3]) {
    double t[3][3];
    t[0][0] = v[2][2] * v[1][1] - v[2][1] * v[1][2];
    t[0][1] = v[2][1] * v[0][2] - v[2][2] * v[0][1];
    t[0][2] = v[1][2] * v[0][1] - v[1][1] * v[0][2];
    double d = v[0][0] * t[0][0] + v[1][0] * t[0][1] + v[2][0] * t[0][2];
    if (d == 0) return false;
    d = 1. / d;
    t[0][0] *= d; t[0][1] *= d; t[0][2] *= d;
    t[1][0] = d * (v[2][0] * v[1][2] - v[2][2] * v[1][0]);
    t[1][1] = d * (v[2][2] * v[0][0] - v[2][0] * v[0][2]);
    t[1][2] = d * (v[1][0] * v[0][2] - v[1][2] * v[0][0]);
    t[2][0] = d * (v[2][1] * v[1][0] - v[2][0] * v[1][1]);
    t[2][1] = d * (v[2][0] * v[0][1] - v[2][1] * v[0][0]);
    t[2][2] = d * (v[1][1] * v[0][0] - v[1][0] * v[0][1]);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) im[a][b] = t[a][b];
    return true;
}
__device__ __forceinline__ bool sd_map_dir(double res[3], const double m[3][3], const double inp[3]) {
    double t[3];
    for (int a = 0; a < 3; a++) t[a] = m[a][0] * inp[0] + m[a][1] * inp[1] + m[a][2] * inp[2];
    if (bsdf_normalize(t) == 0) return false;
    res[0] = t[0]; res[1] = t[1]; res[2] = t[2];
    return true;
}

// ccy2scolor(&c_dfcolor, y) (common/ccyrgb.c:51-65 -> ccy2rgb): the equal-energy white of luminance y through the float
// chromaticity (1/3, 1/3) and the float xyz2rgbmat
__device__ __forceinline__ void bsdf_gray(float col[3], double y) {
    col[0] = (float)y; col[1] = (float)y; col[2] = (float)(y * 0.9999998807907104);
}

}  // namespace rb
