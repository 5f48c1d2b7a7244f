// rb_bins.cuh -- the bin functions of the daylight-coefficient workflow as
// native code (host + device).  Each function restates a .cal file that the
// reference interprets once per contributing ray (rt/rcontrib.c:301-304):
//   reinhartb.cal  rbin    src/radiance/util/reinhartb.cal:15-51
//   reinhart.cal   rbin    src/radiance/cal/cal/reinhart.cal:12-35
//   klems_full.cal kbin    src/radiance/util/klems_full.cal:10-49
//   klems_half / klems_quarter  (same form, other row tables)
//   uniform hemisphere     util/rfluxmtx.c:465-472  if(-Dx*nx-Dy*ny-Dz*nz,0,-1)
// calcomp semantics kept literally: if(c,a,b) is c>0 ? a : b
// (common/calfunc.c:406-413), select() is 1-based, floor() is libm floor.
// D is the direction of the ray that hit the tracked modifier (rt/func.c:443+).
#pragma once
#include <cmath>
#include "rb_device.cuh"

#ifdef __CUDACC__
#define RB_HD __host__ __device__
#else
#define RB_HD
#endif

namespace rb {

RB_HD inline int rb_tnaz(int r) {            // select(r, 30,30,24,24,18,12,6), r = 1..7
    switch (r) {
    case 1: case 2: return 30;
    case 3: case 4: return 24;
    case 5: return 18;
    case 6: return 12;
    case 7: return 6;
    }
    return 0;       // select() out of range is an error in calcomp; unreachable for alt in (0,90]
}
// rnaz(r) = if(r-(7*MF-.5), 1, MF*tnaz(floor((r+.5)/MF) + 1))
RB_HD inline int rb_rnaz(int r, int mf) {
    if (r - (7 * mf - .5) > 0) return 1;
    return mf * rb_tnaz((int)floor((r + .5) / mf) + 1);
}
// raccum(r) = if(r-.5, rnaz(r-1) + raccum(r-1), 0)
RB_HD inline int rb_raccum(int r, int mf) {
    int s = 0;
    for (int k = 0; k < r; k++) s += rb_rnaz(k, mf);
    return s;
}
RB_HD inline int rb_reinhart_nbins(int mf) { return rb_raccum(7 * mf + 1, mf); }   // 144*MF^2 + 1

// common row/azimuth logic of both Reinhart files; alt, azi in degrees
RB_HD inline int rb_reinhart_patch(double alt, double azi, int mf) {
    const double alpha = 90. / (mf * 7 + .5);
    int row = (int)floor(alt / alpha);
    int nz = rb_rnaz(row, mf);
    double inc = 360. / nz;
    int azn = (359.9999 - .5 * inc - azi > 0) ? (int)floor((azi + .5 * inc) / inc) : 0;
    return rb_raccum(row, mf) + azn;
}

RB_HD inline double rb_Asin_deg(double x) {   // Asin(x)/DEGREE
    const double DEGREE = RB_PI / 180;
    double a = (x - 1 > 0) ? RB_PI / 2 : ((-1 - x > 0) ? -RB_PI / 2 : asin(x));
    return a / DEGREE;
}
RB_HD inline double rb_Acos_deg(double x) {
    const double DEGREE = RB_PI / 180;
    double a = (x - 1 > 0) ? 0 : ((-1 - x > 0) ? RB_PI : acos(x));
    return a / DEGREE;
}
RB_HD inline double rb_Atan2_deg(double y, double x) {   // posangle(atan2(y,x))/DEGREE
    const double DEGREE = RB_PI / 180;
    double a = atan2(y, x);
    if (-a > 0) a = a + 2 * RB_PI;
    return a / DEGREE;
}

// Klems row tables: upper polar bound and azimuth count per row
RB_HD inline int rb_klems(double pol, double azi, int fn) {
    if (pol - 90 > 0) return -1;
    // full: kpola {5,15,...,75,90} knaz {1,8,16,20,24,24,24,16,12}
    // half (klems_half.cal:15-16) and quarter (klems_quarter.cal:15-16) tables below
    const double pf[9] = {5, 15, 25, 35, 45, 55, 65, 75, 90};
    const int nf[9] = {1, 8, 16, 20, 24, 24, 24, 16, 12};
    const double ph[7] = {6.5, 19.5, 32.5, 45.5, 58.5, 71.5, 90};
    const int nh[7] = {1, 8, 12, 16, 20, 12, 8};
    const double pq[5] = {9, 27, 45, 63, 90};
    const int nq[5] = {1, 8, 12, 12, 8};
    const double* kp = fn == BIN_KLEMS_FULL ? pf : fn == BIN_KLEMS_HALF ? ph : pq;
    const int* kn = fn == BIN_KLEMS_FULL ? nf : fn == BIN_KLEMS_HALF ? nh : nq;
    int nrows = fn == BIN_KLEMS_FULL ? 9 : fn == BIN_KLEMS_HALF ? 7 : 5;
    // kfindrow(r,pol): first r in 1..nrows with !(pol - kpola(r) > 0); r = nrows if none
    int r = 1, acc = 0;
    while (r < nrows && pol - kp[r - 1] > 0) { acc += kn[r - 1]; r++; }
    double inc = 360. / kn[r - 1];
    int azn = ((360 - .5 * inc) - azi > 0) ? (int)floor((azi + .5 * inc) / inc) : 0;
    return acc + azn;
}
RB_HD inline int rb_klems_nbins(int fn) { return fn == BIN_KLEMS_FULL ? 145 : fn == BIN_KLEMS_HALF ? 77 : 41; }

// Evaluate the bin of direction D for one tracked modifier.  Returns the
// (double) value the .cal expression would; caller applies the reference's
// "<= -.5 ignore, (int)(v+.5)" rule (rt/rcontrib.c:303-306).
RB_HD inline double rb_eval_bin(const DBinSpec& b, const double D[3]) {
    switch (b.fn) {
    case BIN_CONST:
        return (double)b.cbin;
    case BIN_HEMI: {
        double v = -D[0] * b.n[0] - D[1] * b.n[1] - D[2] * b.n[2];
        return v > 0 ? 0. : -1.;
    }
    case BIN_REINHARTB: {
        const double* N = b.n; const double* U = b.u;
        double dz = -D[0] * N[0] - D[1] * N[1] - D[2] * N[2];
        double rx = -b.rhs * (D[0] * (U[1] * N[2] - U[2] * N[1]) + D[1] * (U[2] * N[0] - U[0] * N[2]) +
                              D[2] * (U[0] * N[1] - U[1] * N[0]));
        double ry = D[0] * U[0] + D[1] * U[1] + D[2] * U[2] + dz * (N[0] * U[0] + N[1] * U[1] + N[2] * U[2]);
        double alt = rb_Asin_deg(dz);
        double azi = rb_Atan2_deg(rx, ry);
        if (!(alt > 0)) return -1.;
        return (double)rb_reinhart_patch(alt, azi, b.mf);
    }
    case BIN_REINHART: {
        double alt = rb_Asin_deg(D[2]);
        double azi = rb_Atan2_deg(D[0], D[1]);
        if (-alt > 0) return 0.;
        return (double)(rb_reinhart_patch(alt, azi, b.mf) + 1);
    }
    case BIN_KLEMS_FULL: case BIN_KLEMS_HALF: case BIN_KLEMS_QUARTER: {
        const double* N = b.n; const double* U = b.u;
        double pol = rb_Acos_deg(-D[0] * N[0] - D[1] * N[1] - D[2] * N[2]);
        double y = -D[0] * U[0] - D[1] * U[1] - D[2] * U[2] +
                   (N[0] * D[0] + N[1] * D[1] + N[2] * D[2]) * (N[0] * U[0] + N[1] * U[1] + N[2] * U[2]);
        double x = -b.rhs * (D[0] * (U[1] * N[2] - U[2] * N[1]) + D[1] * (U[2] * N[0] - U[0] * N[2]) +
                             D[2] * (U[0] * N[1] - U[1] * N[0]));
        return (double)rb_klems(pol, rb_Atan2_deg(y, x), b.fn);
    }
    case BIN_SHIRCHIU: {          // util/disk2square.cal:62-76 scbin, :43-60 disk -> square (SCdim in b.mf)
        const double* N = b.n; const double* U = b.u;
        const double PI_ = 3.14159265358979323846;
        double dz = -D[0] * N[0] - D[1] * N[1] - D[2] * N[2];
        double rx = -b.rhs * (D[0] * (U[1] * N[2] - U[2] * N[1]) + D[1] * (U[2] * N[0] - U[0] * N[2]) +
                              D[2] * (U[0] * N[1] - U[1] * N[0]));
        double ry = D[0] * U[0] + D[1] * U[1] + D[2] * U[2] + dz * (N[0] * U[0] + N[1] * U[1] + N[2] * U[2]);
        double den2 = rx * rx + ry * ry;
        double radf = (den2 - 1e-7 > 0) ? sqrt((1 - dz * dz) / den2) : 0.;
        double dx = rx * radf, dy = -ry * radf;
        double r = sqrt(dx * dx + dy * dy);
        double phi = atan2(dy, dx);
        if (-phi - PI_ / 4 > 0) phi += 2 * PI_;                    // norm_radians
        int rgn = (int)floor((phi + PI_ / 4) / (PI_ / 2)) + 1;     // select() index 1..5
        double a, bb;
        switch (rgn) {
        case 1: a = r; bb = phi * r / (PI_ / 4); break;
        case 2: a = (PI_ / 2 - phi) * r / (PI_ / 4); bb = r; break;
        case 3: a = -r; bb = (PI_ - phi) * r / (PI_ / 4); break;
        case 4: a = (phi - 3 * PI_ / 2) * r / (PI_ / 4); bb = -r; break;
        default: a = r; bb = -r; break;
        }
        double sx = (a + 1) / 2, sy = (bb + 1) / 2;
        if (!(dz > 0)) return -1.;
        return floor(sx * b.mf) * b.mf + floor(sy * b.mf);
    }
    }
    return -1.;
}

}  // namespace rb
