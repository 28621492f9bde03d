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

// e and (optionally) de/dv per channel, given c = max(cos_i, 0) and the metric weight w of the direction pair
BBMCU_D float loss_term_g(int metric, float c, float w, const Spec<float>& v, const Spec<float>& r, Spec<float>* dv)
{
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

BBMCU_D float loss_term(int metric, f3 in, f3 out, const Spec<float>& v, const Spec<float>& r, Spec<float>* dv)
{ return loss_term_g(metric, fmaxf(in.z, 0.0f), metric_weight(metric, in, out), v, r, dv); }

// ---- fit-parameter layout of a model: which attribute floats are fit parameters ----------------------
// Default: every attribute float, in order.  Bagher keeps 15 Dependent floats (K, Lambda, c, theta0, k)
// between its albedo and the fit parameters alpha, p, eta (ndf/sgd.h:198-202, bsdf_attr_flag.h:23,28).
template<class M> struct FitMap
{
  static constexpr int NFIT = M::NA;
  BBMCU_HD static constexpr int attr_of(int k) { return k; }
};
template<> struct FitMap<MerlModel> { static constexpr int NFIT = 0; BBMCU_HD static constexpr int attr_of(int k) { return k; } };   // measured data: nothing to fit
using BagherModel = Microfacet<NdfSGD, GUncorrelated, FresnelBagher, 2, true>;
template<> struct FitMap<BagherModel>
{
  static constexpr int NFIT = 15;
  BBMCU_HD static constexpr int attr_of(int k) { return k < 3 ? k : k + 15; }
};

// number of fit parameters of a model that are NOT its leading RGB scale
template<class M> struct NonLinear { static constexpr int N = FitMap<M>::NFIT - (M::SCALE >= 0 ? 3 : 0); };

// ---- direction-only precomputation ("geometry") -------------------------------------------------------------------
// A model may declare `struct Geom`, `geom(in, out)` and `eval_unscaled_g<T>(geom, in, out, a, component)`: everything
// that depends on the direction pair alone (half vector, dots).  The batched loss kernels compute it once per sample,
// outside their loop over parameter sets.
template<class...> struct VoidT { using type = void; };
// models whose value-only loss passes evaluate the all-float form of their templates (Dual<0>) instead of eval<float> with the
// reference's double spots: the He family, whose Taylor series (a double exp and float <-> double conversions per term and
// channel) is two thirds of the configs[4] sweep.  Same value as the gradient kernels return (they carry it as the value part
// of a dual number), within ~1e-5 of eval<float> per sample - the loss contract is 1e-4 on the total.
template<class M, class = void> struct QuickLossValue { static constexpr bool value = false; };
template<class M> struct QuickLossValue<M, typename std::enable_if<M::kQuickLossValue>::type> { static constexpr bool value = true; };
template<class M, class = void> struct GeomOf
{
  struct type {};
  BBMCU_D static type make(f3, f3) { return type(); }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(const type&, f3 in, f3 out, const T* a, int component) { return M::template eval_unscaled<T>(in, out, a, component); }
  BBMCU_D static Spec<float> eval(const type&, f3 in, f3 out, const float* a, int component)
  {
    if constexpr (QuickLossValue<M>::value)
    {
      Dual<0> ad[M::NA];
#pragma unroll
      for(int i=0; i < M::NA; ++i) ad[i] = Dual<0>(a[i]);
      const Spec<Dual<0>> u = M::template eval_unscaled<Dual<0>>(in, out, ad, component);
      Spec<float> r(u.r.v, u.g.v, u.b.v);
      if(M::SCALE >= 0) r = r * load_spec(a + (M::SCALE >= 0 ? M::SCALE : 0));
      return r;
    }
    else return M::template eval<float>(in, out, a, component);
  }
};
template<class M> struct GeomOf<M, typename VoidT<typename M::Geom>::type>
{
  using type = typename M::Geom;
  BBMCU_D static type make(f3 in, f3 out) { return M::geom(in, out); }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(const type& g, f3 in, f3 out, const T* a, int component) { return M::template eval_unscaled_g<T>(g, in, out, a, component); }
  BBMCU_D static Spec<float> eval(const type& g, f3 in, f3 out, const float* a, int component)
  { Spec<float> r = M::template eval_unscaled_g<float>(g, in, out, a, component); if(M::SCALE >= 0) r = r * load_spec(a + (M::SCALE >= 0 ? M::SCALE : 0)); return r; }
};
// models whose unscaled value has three identical channels (scalar D, G and F): the jacobian is one column
template<class M, class = void> struct GrayUnscaled { static constexpr bool value = false; };
template<class M> struct GrayUnscaled<M, typename std::enable_if<M::kGrayUnscaled>::type> { static constexpr bool value = true; };

// microfacet lobes whose jacobian separates into an NDF block and a Fresnel block (Microfacet::unscaled_gray_jacobian):
// scaled, gray before the scale, every NDF and Fresnel parameter a fit parameter in attribute order
template<class M, class = void> struct SeparableJac { static constexpr bool value = false; };
template<class M> struct SeparableJac<M, typename std::enable_if<(M::kNdfParams + M::kFresnelParams > 0)>::type>
{
  static constexpr bool value = GrayUnscaled<M>::value && (M::SCALE == 0) && FitMap<M>::NFIT == M::NA && FitMap<M>::attr_of(M::NA - 1) == M::NA - 1
                             && NonLinear<M>::N == M::kNdfParams + M::kFresnelParams;
};

// value and parameter jacobian of one lobe.  Scale parameters (a leading RGB attribute) touch one channel each and are
// kept as the three unscaled values `us`; the NL non-linear parameters carry a full RGB column.
template<class M> struct LobeJac
{
  static constexpr int NL = NonLinear<M>::N;
  static constexpr bool SCALED = (M::SCALE >= 0);
  Spec<float> v;
  float us[3];
  float jr[NL > 0 ? NL : 1], jg[NL > 0 ? NL : 1], jb[NL > 0 ? NL : 1];
  // d e / d theta_k for this lobe's NFIT parameters, given d e / d v
  BBMCU_D void combine(const Spec<float>& dv, const float* a, float* grad) const
  {
    constexpr int S = SCALED ? 3 : 0;
    if constexpr (FitMap<M>::NFIT == 0) { (void)dv; (void)a; (void)grad; return; }
    if(SCALED) { grad[0] = dv.r*us[0]; grad[1] = dv.g*us[1]; grad[2] = dv.b*us[2]; }
    if constexpr (NL > 0)
    {
      if constexpr (SCALED && GrayUnscaled<M>::value)
      {
        const float sdv = (dv.r*a[0] + dv.g*a[1]) + dv.b*a[2];
#pragma unroll
        for(int j=0; j < NL; ++j) grad[S + j] = sdv*jr[j];               // jr holds d u / d theta (unscaled) here
      }
      else
      {
#pragma unroll
        for(int j=0; j < NL; ++j) grad[S + j] = (dv.r*jr[j] + dv.g*jg[j]) + dv.b*jb[j];
      }
    }
  }
};

template<class M>
BBMCU_D LobeJac<M> lobe_jacobian(const typename GeomOf<M>::type& geom, const float* a, f3 in, f3 out, int component)
{
  constexpr int NL = NonLinear<M>::N;
  constexpr int NFIT = FitMap<M>::NFIT;
  constexpr bool SCALED = (M::SCALE >= 0);
  static_assert(!SCALED || M::SCALE == 0, "leading scale expected at offset 0");
  LobeJac<M> J;
  if constexpr (NFIT == 0)
  {
    J.v = GeomOf<M>::template eval_unscaled<float>(geom, in, out, a, component);
    J.us[0] = J.us[1] = J.us[2] = 0.0f;
  }
  else if constexpr (NL == 0)
  {
    Spec<float> u = GeomOf<M>::template eval_unscaled<float>(geom, in, out, a, component);
    J.us[0] = u.r; J.us[1] = u.g; J.us[2] = u.b;
    J.v = Spec<float>(a[0]*u.r, a[1]*u.g, a[2]*u.b);
  }
  else if constexpr (SeparableJac<M>::value)
  {
    const Dual<NL> u = M::template unscaled_gray_jacobian<NL>(geom, in, out, a, component);
    J.us[0] = J.us[1] = J.us[2] = u.v;
#pragma unroll
    for(int j=0; j < NL; ++j) J.jr[j] = u.d[j];
    J.v = Spec<float>(a[0]*u.v, a[1]*u.v, a[2]*u.v);
  }
  else
  {
    using D = Dual<NL>;
    D ad[M::NA];
#pragma unroll
    for(int i=0; i < M::NA; ++i) ad[i] = D(a[i]);
#pragma unroll
    for(int k = (SCALED ? 3 : 0); k < NFIT; ++k) ad[FitMap<M>::attr_of(k)].d[k - (SCALED ? 3 : 0)] = 1.0f;
    Spec<D> u = GeomOf<M>::template eval_unscaled<D>(geom, in, out, ad, component);
    const float sr = SCALED ? a[0] : 1.0f, sg = SCALED ? a[1] : 1.0f, sb = SCALED ? a[2] : 1.0f;
    J.us[0] = u.r.v; J.us[1] = u.g.v; J.us[2] = u.b.v;
    if constexpr (SCALED && GrayUnscaled<M>::value)
    {
#pragma unroll
      for(int j=0; j < NL; ++j) J.jr[j] = u.r.d[j];
    }
    else
    {
#pragma unroll
      for(int j=0; j < NL; ++j) { J.jr[j] = sr*u.r.d[j]; J.jg[j] = sg*u.g.d[j]; J.jb[j] = sb*u.b.d[j]; }
    }
    J.v = Spec<float>(sr*u.r.v, sg*u.g.v, sb*u.b.v);
  }
  return J;
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
        LobeJac<M> J = lobe_jacobian<M>(GeomOf<M>::make(in, out), a, in, out, component);
        J.combine(dv, a, grad + base);
        base += NFitOf<M>::N;
      });
    }
  }
  return e;
}

// per-sample quantities of the metric that do not depend on the fitted model
struct TermGeom { float c, w; };
BBMCU_D TermGeom term_geom(int metric, f3 in, f3 out) { TermGeom t; t.c = fmaxf(in.z, 0.0f); t.w = metric_weight(metric, in, out); return t; }

// ---- compile-time one- and two-lobe BSDFs: single pass, everything in registers --------------------------
template<class M>
struct LossSingle
{
  static constexpr int P = NFitOf<M>::N;
  struct Geom { typename GeomOf<M>::type g; TermGeom t; };
  BBMCU_D static Geom geom(int metric, f3 in, f3 out) { Geom G; G.g = GeomOf<M>::make(in, out); G.t = term_geom(metric, in, out); return G; }
  BBMCU_D static float sample(const float* attrs, int metric, int component, const Geom& G, f3 in, f3 out, const Spec<float>& ref, float (&grad)[P], bool want_grad)
  {
    if(!want_grad) { Spec<float> v = GeomOf<M>::eval(G.g, in, out, attrs, component); return loss_term_g(metric, G.t.c, G.t.w, v, ref, nullptr); }
    LobeJac<M> J = lobe_jacobian<M>(G.g, attrs, in, out, component);
    Spec<float> dv;
    float e = loss_term_g(metric, G.t.c, G.t.w, J.v, ref, &dv);
    J.combine(dv, attrs, grad);
    return e;
  }
};

// Aggregate(M0, M1): the shape of every entry of the reference's fits/*.fit (Lambertian + specular lobe)
template<class M0, class M1>
struct LossPair
{
  static constexpr int P0 = NFitOf<M0>::N, P = P0 + NFitOf<M1>::N;
  struct Geom { typename GeomOf<M0>::type g0; typename GeomOf<M1>::type g1; TermGeom t; };
  BBMCU_D static Geom geom(int metric, f3 in, f3 out) { Geom G; G.g0 = GeomOf<M0>::make(in, out); G.g1 = GeomOf<M1>::make(in, out); G.t = term_geom(metric, in, out); return G; }
  BBMCU_D static float sample(const float* attrs, int metric, int component, const Geom& G, f3 in, f3 out, const Spec<float>& ref, float (&grad)[P], bool want_grad)
  {
    const float* a0 = attrs; const float* a1 = attrs + M0::NA;
    if(!want_grad)
    {
      Spec<float> v = (Spec<float>(0.0f) + GeomOf<M0>::eval(G.g0, in, out, a0, component)) + GeomOf<M1>::eval(G.g1, in, out, a1, component);
      return loss_term_g(metric, G.t.c, G.t.w, v, ref, nullptr);
    }
    LobeJac<M0> J0 = lobe_jacobian<M0>(G.g0, a0, in, out, component);
    LobeJac<M1> J1 = lobe_jacobian<M1>(G.g1, a1, in, out, component);
    Spec<float> dv;
    float e = loss_term_g(metric, G.t.c, G.t.w, J0.v + J1.v, ref, &dv);
    J0.combine(dv, a0, grad);
    J1.combine(dv, a1, grad + P0);
    return e;
  }
};

} // namespace bbmcu
