// rbk_hostmath.cpp - compiles the PRODUCT's device arithmetic (rbk_math.cuh / rbk_step.cuh) for the
// host, so that the CPU test-suite can check it against the oracle without a GPU
// (tests/test_device_math_host.py).  Not part of librbk.so and not a fallback: nothing in the
// product calls this library.
#include "rbk_step.cuh"

using namespace rbk;

extern "C" {

void rbkh_jacobi(double u, double m, double* sn, double* cn, double* dn) { jacobiSnCnDn(u, m, *sn, *cn, *dn); }
double rbkh_rf(double x, double y, double z) { return carlsonRF(x, y, z); }
double rbkh_rj(double x, double y, double z, double p) { return carlsonRJ(x, y, z, p); }
double rbkh_rc(double x, double y) { return carlsonRC(x, y); }

void rbkh_exact_rotation(double dt, const double* I, double* q, double* pi, int elliptic_only) {
    d3 Iv = {I[0], I[1], I[2]}, inv = {1.0/I[0], 1.0/I[1], 1.0/I[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pp = {pi[0], pi[1], pi[2], pi[3]};
    if (elliptic_only) exactRotationElliptic(dt, Iv, inv, qq, pp);
    else exactRotation(dt, inv, qq, pp);
    q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z;
    pi[0] = pp.w; pi[1] = pp.x; pi[2] = pp.y; pi[3] = pp.z;
}

// Series fast path alone with a chosen order; returns 1 when its truncation check passed.
int rbkh_exact_series(int order, double dt, const double* I, double* q, double* pi) {
    d3 inv = {1.0/I[0], 1.0/I[1], 1.0/I[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pp = {pi[0], pi[1], pi[2], pi[3]};
    bool ok = false;
    switch (order) {
    case 6: ok = exactRotationSeries<6>(dt, inv, qq, pp) == 0.0; break;
    case 8: ok = exactRotationSeries<8>(dt, inv, qq, pp) == 0.0; break;
    case 10: ok = exactRotationSeries<10>(dt, inv, qq, pp) == 0.0; break;
    case 11: ok = exactRotationSeries<11>(dt, inv, qq, pp) == 0.0; break;
    case 12: ok = exactRotationSeries<12>(dt, inv, qq, pp) == 0.0; break;
    case 13: ok = exactRotationSeries<13>(dt, inv, qq, pp) == 0.0; break;
    case 14: ok = exactRotationSeries<14>(dt, inv, qq, pp) == 0.0; break;
    case 16: ok = exactRotationSeries<16>(dt, inv, qq, pp) == 0.0; break;
    case 20: ok = exactRotationSeries<20>(dt, inv, qq, pp) == 0.0; break;
    default: return -1;
    }
    q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z;
    pi[0] = pp.w; pi[1] = pp.x; pi[2] = pp.y; pi[3] = pp.z;
    return ok ? 1 : 0;
}

// The fast path's overshoot of its truncation bound at the production order (0 = converged): what the retry path
// turns into a number of sub-steps.  q, pi are not modified.
double rbkh_series_excess(double dt, const double* I, const double* q, const double* pi) {
    d3 inv = {1.0/I[0], 1.0/I[1], 1.0/I[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pp = {pi[0], pi[1], pi[2], pi[3]};
    return exactRotationSeries<kSeriesOrder>(dt, inv, qq, pp);
}

// The ladder entry point of the hot kernels: rung 0, 1, 2 = order 11, 13, 16; returns the flags (bit 0: check failed at this
// rung and the retry path ran, bit 1: the check would have failed one rung lower).
unsigned rbkh_exact_ladder(int rung, double dt, const double* I, double* q, double* pi) {
    d3 inv = {1.0/I[0], 1.0/I[1], 1.0/I[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pp = {pi[0], pi[1], pi[2], pi[3]};
    unsigned flags = 0;
    exactRotationLadder(rung, dt, inv, qq, pp, flags);
    q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z;
    pi[0] = pp.w; pi[1] = pp.x; pi[2] = pp.y; pi[3] = pp.z;
    return flags;
}

void rbkh_nosquish(double dt, int n, const double* invI, double* q, double* pi) {
    d3 inv = {invI[0], invI[1], invI[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pp = {pi[0], pi[1], pi[2], pi[3]};
    noSquish(dt, n, inv, qq, pp);
    q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z;
    pi[0] = pp.w; pi[1] = pp.x; pi[2] = pp.y; pi[3] = pp.z;
}

// One body through Part 1 (mode 0 or n) exactly as the kernel's thread-per-body phase does it.
void rbkh_body_part1(int mode, double dt, const double* F, const double* tau, double invm, const double* I,
                     const double* invI, double* r, double* p, double* q, double* pi) {
    d3 rr = {r[0], r[1], r[2]}, pp = {p[0], p[1], p[2]};
    d4 qq = {q[0], q[1], q[2], q[3]}, pq = {pi[0], pi[1], pi[2], pi[3]};
    d3 Fv = {F[0], F[1], F[2]}, tv = {tau[0], tau[1], tau[2]}, Iv = {I[0], I[1], I[2]}, inv = {invI[0], invI[1], invI[2]};
    (void) Iv;
    if (mode == 0) bodyPart1<true>(dt, 0, Fv, tv, invm, inv, rr, pp, qq, pq);
    else bodyPart1<false>(dt, mode, Fv, tv, invm, inv, rr, pp, qq, pq);
    r[0] = rr.x; r[1] = rr.y; r[2] = rr.z; p[0] = pp.x; p[1] = pp.y; p[2] = pp.z;
    q[0] = qq.w; q[1] = qq.x; q[2] = qq.y; q[3] = qq.z; pi[0] = pq.w; pi[1] = pq.x; pi[2] = pq.y; pi[3] = pq.z;
}

} // extern "C"
