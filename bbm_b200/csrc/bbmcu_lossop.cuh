// Per-sample fitting error of the six metrics and its analytic parameter gradient.
//
// Error functors restate include/loss/cosine_weighted_l2.h:25-34,96-105,166-176 and
// include/loss/cosine_weighted_log.h:32-43,101-112,170-181 including their float/double mix:
//   L2 family : hsum(pow((v - r) * max(cos_i, 0), 2.0)) * w       -> squares and sum in double
//   log family: hsum(pow(log(1 + v c) - log(1 + r c), 2.0)) * w   -> logs in float, squares in double
//   w = sin_i sin_o (ngan, standardLog), sin_i (low*), max(cos_o,0) sin_i sin_o (bieron*), applied
//   left to right in double; the value is rounded to float once (the functor returns Value).
//
// The gradient is new capability (the reference has none):
//   d e / d theta_j = sum_c  de/dv_c * dv_c/dtheta_j,   v = sum over lobes of scale_l * u_l(theta_l)
// with du_l/dtheta from forward-mode dual numbers through the very same model code
// (eval_unscaled<Dual<N>>) and d/dscale in closed form.
#pragma once
#include "bbmcu_bsdf.cuh"

namespace bbmcu {

constexpr int kMaxParams = 32;      // fit parameters per BSDF (Aggregate(Lambertian, Bagher) = 18)

enum : int { METRIC_NGAN_L2 = 0, METRIC_LOW_L2, METRIC_BIERON_L2, METRIC_LOW_LOG, METRIC_BIERON_LOG, METRIC_STANDARD_LOG };

// w(in, out): sin_i sin_o (ngan, standardLog), sin_i (low*), max(cos_o, 0) sin_i sin_o (bieron*).  The reference forms
// e * w left to right in double and rounds once; every factor is non-negative, so float products are within 3e-7.
BBMCU_D float metric_weight(int metric, f3 in, f3 out)
{
  switch(metric) {
    case METRIC_NGAN_L2: case METRIC_STANDARD_LOG: return q_sinTheta(in) * q_sinTheta(out);
    case METRIC_LOW_L2: case METRIC_LOW_LOG:       return q_sinTheta(in);
    default:                                       return fmaxf(out.z, 0.0f) * q_sinTheta(in) * q_sinTheta(out);
  }
}

// e and (optionally) de/dv per channel
BBMCU_D float loss_term(int metric, f3 in, f3 out, const Spec<float>& v, const Spec<float>& r, Spec<float>* dv)
{
  float c = fmaxf(in.z, 0.0f);
  float w = metric_weight(metric, in, out);
  if(metric <= METRIC_BIERON_L2)
  {
    float tr = (v.r - r.r)*c, tg = (v.g - r.g)*c, tb = (v.b - r.b)*c;
    float s = (tr*tr + tg*tg) + tb*tb;
    if(dv) { float k = 2.0f * w * c; *dv = Spec<float>(k*tr, k*tg, k*tb); }
    return s * w;
  }
  float ar = 1.0f + v.r*c, ag = 1.0f + v.g*c, ab = 1.0f + v.b*c;
  float dr = logf(ar) - logf(1.0f + r.r*c), dg = logf(ag) - logf(1.0f + r.g*c), db = logf(ab) - logf(1.0f + r.b*c);
  float s = (dr*dr + dg*dg) + db*db;
  if(dv) { float k = 2.0f * w * c; *dv = Spec<float>(k*q_div(dr, ar), k*q_div(dg, ag), k*q_div(db, ab)); }
  return s * w;
}

// ---- fit-parameter layout of a model: which attribute floats are fit parameters ----------------------
// Default: every attribute float, in order.  Bagher keeps 15 Dependent floats (K, Lambda, c, theta0, k)
// between its albedo and the fit parameters alpha, p, eta (ndf/sgd.h:198-202, bsdf_attr_flag.h:23,28).
template<class M> struct FitMap
{
  static constexpr int NFIT = M::NA;
  BBMCU_HD static constexpr int attr_of(int k) { return k; }
};
using BagherModel = Microfacet<NdfSGD, GUncorrelated, FresnelBagher, 2, true>;
template<> struct FitMap<BagherModel>
{
  static constexpr int NFIT = 15;
  BBMCU_HD static constexpr int attr_of(int k) { return k < 3 ? k : k + 15; }
};

// number of fit parameters of a model that are NOT its leading RGB scale
template<class M> struct NonLinear { static constexpr int N = FitMap<M>::NFIT - (M::SCALE >= 0 ? 3 : 0); };

// One lobe: value v (added into `v`) and, through `emit(k, dv_r, dv_g, dv_b)`, dv/dtheta_k for each of the
// lobe's fit parameters k = 0 .. NFIT-1 (forward enumeration order).
template<class M, class Emit>
BBMCU_D Spec<float> lobe_value_and_jacobian(const float* a, f3 in, f3 out, int component, Emit&& emit)
{
  constexpr int NL = NonLinear<M>::N;
  constexpr int NFIT = FitMap<M>::NFIT;
  constexpr bool SCALED = (M::SCALE >= 0);
  static_assert(!SCALED || M::SCALE == 0, "leading scale expected at offset 0");
  if constexpr (NL == 0)
  {
    Spec<float> u = M::template eval_unscaled<float>(in, out, a, component);
    emit(0, u.r, 0.0f, 0.0f); emit(1, 0.0f, u.g, 0.0f); emit(2, 0.0f, 0.0f, u.b);
    return Spec<float>(a[0]*u.r, a[1]*u.g, a[2]*u.b);
  }
  else
  {
    using D = Dual<NL>;
    D ad[M::NA];
#pragma unroll
    for(int i=0; i < M::NA; ++i) ad[i] = D(a[i]);
#pragma unroll
    for(int k = (SCALED ? 3 : 0); k < NFIT; ++k) ad[FitMap<M>::attr_of(k)].d[k - (SCALED ? 3 : 0)] = 1.0f;
    Spec<D> u = M::template eval_unscaled<D>(in, out, ad, component);
    float sr = SCALED ? a[0] : 1.0f, sg = SCALED ? a[1] : 1.0f, sb = SCALED ? a[2] : 1.0f;
    if constexpr (SCALED) { emit(0, u.r.v, 0.0f, 0.0f); emit(1, 0.0f, u.g.v, 0.0f); emit(2, 0.0f, 0.0f, u.b.v); }
#pragma unroll
    for(int j=0; j < NL; ++j) emit((SCALED ? 3 : 0) + j, sr*u.r.d[j], sg*u.g.d[j], sb*u.b.d[j]);
    return Spec<float>(sr*u.r.v, sg*u.g.v, sb*u.b.v);
  }
}

template<class M> struct NFitOf { static constexpr int N = FitMap<M>::NFIT; };

// ---- run-time lobe list: value pass, then one jacobian pass per lobe ----------------------------------
// returns e; if grad != nullptr writes de/dtheta for all fit parameters of the BSDF (forward order)
BBMCU_D float loss_sample_generic(const BsdfDesc& b, int metric, int component, f3 in, f3 out, const Spec<float>& ref, float* grad)
{
  Spec<float> v = BsdfGeneric::eval(b, in, out, component);
  Spec<float> dv;
  float e = loss_term(metric, in, out, v, ref, grad ? &dv : nullptr);
  if(grad)
  {
    int base = 0;
    for(int l=0; l < b.n_lobes; ++l)
    {
      const float* a = b.attrs + b.offset[l];
      dispatch_model(b.model[l], [&](auto* tag) {
        using M = typename std::remove_pointer<decltype(tag)>::type;
        lobe_value_and_jacobian<M>(a, in, out, component, [&](int k, float jr, float jg, float jb) { grad[base + k] = dv.r*jr + dv.g*jg + dv.b*jb; });
        base += NFitOf<M>::N;
      });
    }
  }
  return e;
}

// ---- compile-time one- and two-lobe BSDFs: single pass, everything in registers --------------------------
template<class M>
struct LossSingle
{
  static constexpr int P = NFitOf<M>::N;
  BBMCU_D static float sample(const float* attrs, int metric, int component, f3 in, f3 out, const Spec<float>& ref, float (&grad)[P], bool want_grad)
  {
    if(!want_grad) { Spec<float> v = M::template eval<float>(in, out, attrs, component); return loss_term(metric, in, out, v, ref, nullptr); }
    float jr[P], jg[P], jb[P];
    Spec<float> v = lobe_value_and_jacobian<M>(attrs, in, out, component, [&](int k, float r, float g, float b) { jr[k] = r; jg[k] = g; jb[k] = b; });
    Spec<float> dv;
    float e = loss_term(metric, in, out, v, ref, &dv);
#pragma unroll
    for(int k=0; k < P; ++k) grad[k] = dv.r*jr[k] + dv.g*jg[k] + dv.b*jb[k];
    return e;
  }
};

// Aggregate(M0, M1): the shape of every entry of the reference's fits/*.fit (Lambertian + specular lobe)
template<class M0, class M1>
struct LossPair
{
  static constexpr int P0 = NFitOf<M0>::N, P = P0 + NFitOf<M1>::N;
  BBMCU_D static float sample(const float* attrs, int metric, int component, f3 in, f3 out, const Spec<float>& ref, float (&grad)[P], bool want_grad)
  {
    const float* a0 = attrs; const float* a1 = attrs + M0::NA;
    if(!want_grad)
    {
      Spec<float> v = (Spec<float>(0.0f) + M0::template eval<float>(in, out, a0, component)) + M1::template eval<float>(in, out, a1, component);
      return loss_term(metric, in, out, v, ref, nullptr);
    }
    float jr[P], jg[P], jb[P];
    Spec<float> v0 = lobe_value_and_jacobian<M0>(a0, in, out, component, [&](int k, float r, float g, float b) { jr[k] = r; jg[k] = g; jb[k] = b; });
    Spec<float> v1 = lobe_value_and_jacobian<M1>(a1, in, out, component, [&](int k, float r, float g, float b) { jr[P0 + k] = r; jg[P0 + k] = g; jb[P0 + k] = b; });
    Spec<float> dv;
    float e = loss_term(metric, in, out, v0 + v1, ref, &dv);
#pragma unroll
    for(int k=0; k < P; ++k) grad[k] = dv.r*jr[k] + dv.g*jg[k] + dv.b*jb[k];
    return e;
  }
};

} // namespace bbmcu
