// nbconst.h -- the constant block shared by host (filled in nb200_set_stencil/_set_collision)
// and device (one __constant__ copy per stencil unit).
#pragma once

#define NB_MAXQ 45

struct NbConst {
    double e[NB_MAXQ][3];   // unscaled directions (e_scaled / scaling)
    double es[NB_MAXQ][3];  // scaled directions (diagnostics)
    double w[NB_MAXQ];
    double inv_w[NB_MAXQ];
    // unique Hermite components, order:
    // H3: xxx xxy xyy yyy | zzz xxz xzz yzz yyz xyz
    // H4: xxxx yyyy xxxy xyyy xxyy | zzzz xzzz xxzz xxxz yzzz yyzz yyyz xxyz xyyz xyzz
    double H3[NB_MAXQ][10];
    double H4[NB_MAXQ][15];
    double cs2;        // unscaled speed of sound squared
    double scaling;
    double tau;
    double gamma, Cv, prandtl;
    int prandtl_set, sutherland_set;
    int D, Q;
};


enum { NB_EQ_BGK = 0, NB_EQ_QUARTIC = 1 };
