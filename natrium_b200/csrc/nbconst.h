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
    double inv_cs2, half_inv_cs2;   // 1/cs2, 1/(2 cs2)
    double inv_c3, inv_c4;          // 1/(6 cs2^3), 1/(24 cs2^4)
    double inv_tau;                 // 1/tau
    double scaling;
    double tau;
    double tau_legacy;   // nu/(dt*cs2_scaled): relaxation parameter of the legacy CollisionModel family
    double gamma, Cv, prandtl;
    int prandtl_set, sutherland_set;
    int has_force, force_type;      // problemDescription.hasExternalForce(), ForceType (ConfigNames.h:114-119)
    double force[3];                // getExternalForce()->getForce()
    double dt;
    int D, Q;
};


// collision kind = template parameter of the kernels: the two equilibria of the collision_advanced BGK
// scheme, and the legacy entropic models
enum { NB_EQ_BGK = 0, NB_EQ_QUARTIC = 1, NB_KIND_KBC = 2, NB_KIND_MRT_ENTROPIC = 3, NB_KIND_REGULARIZED = 4, NB_KIND_MRT = 5 };

// MultipleRelaxationTime tables (host copy, uploaded to the D2Q9 / D3Q19 units)
struct NbMrtStdHost {
    double M[19][19];
    double T[19][19];
    double omega[19];
};

// MRTEntropic D3Q19 moment matrix and inverse (host copy, uploaded to the D3Q19 unit)
struct NbMrtHost {
    double tm[19][19];
    double invm[19][19];
};
