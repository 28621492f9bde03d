/* TEST INFRASTRUCTURE ONLY - never linked into, loaded by or called from the product path (libbbmcu.so).
 *
 * Plain-C restatement of the core of bbm's hot path, written from the reading in SURVEY.md section 8(a); it is the
 * "port" oracle next to the compiled unmodified reference (oracle/_ref, the primary oracle).  Pinned by
 * tests/test_oracle_port.py against the golden vectors the unmodified reference produced (tests/golden) and, when
 * oracle/_ref is built, against the reference itself.  Every function cites the reference lines it follows.
 *
 *   bbmo_merl_index      include/linearizer/merl_linearizer.h:94-123 (+ core/spherical.h:26-46, core/vec_transform.h:77-97,
 *                        core/transform.h:47-54,74-81, core/mat.h:107-116, backbone/native horizontal.h:87-91,106)
 *   bbmo_eval            Lambertian (bsdfmodel/lambertian.h:45-59), CookTorrance (cooktorrance.h:29-34 ->
 *                        microfacet.h:74-102, ndf/beckmann.h:49-66, maskingshadowing/vgroove.h:30-47,
 *                        bbm/fresnel_cook.h:41-56, scaledmodel.h:50-53), GGX (ggx.h:28-33, ndf/ggx.h:50-65,173-189,
 *                        maskingshadowing/uncorrelated.h:30-42)
 *   bbmo_ngan_l2_term    loss/cosine_weighted_l2.h:25-34
 *
 * Float / double mix as the native backbone evaluates it; compile with -ffp-contract=off. */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define PI_F   3.14159265358979323846f
#define TWOPI_F 6.28318530717958647692f
#define HALFPI_F 1.57079632679489661923f
#define EPS_F  1.1920928955078125e-07f

typedef struct { float x, y, z; } v3;
static float dot3(v3 a, v3 b) { return ((0.0f + a.x*b.x) + a.y*b.y) + a.z*b.z; }
static v3 normalize3(v3 v) { float r = 1.0f / sqrtf(dot3(v, v)); v3 o = { v.x*r, v.y*r, v.z*r }; return o; }
static float phi3(v3 v) { float r = atan2f(v.y, v.x); return r < 0.0f ? r + TWOPI_F : r; }
static float theta3(v3 v)
{
  v3 d = v; d.z -= copysignf(1.0f, v.z);
  float n = sqrtf(dot3(d, d));
  double t = 2.0 * asin(0.5 * (double)n);
  return v.z >= 0.0f ? (float)t : (float)((double)PI_F - t);
}
static float clampf(float a, float lo, float hi) { return a < lo ? lo : (hi < a ? hi : a); }

/* direction pair -> MERL bin; 1458000 below the horizon, 0xFFFFFFFF for NaN half vectors */
uint32_t bbmo_merl_index(const float* in, const float* out)
{
  v3 i = { in[0], in[1], in[2] }, o = { out[0], out[1], out[2] };
  if(!(i.z >= 0.0f && o.z >= 0.0f)) return 1458000u;
  v3 s = { i.x + o.x, i.y + o.y, i.z + o.z };
  v3 h = normalize3(s);
  float ph = phi3(h), th = theta3(h);
  float c = cosf(-ph), sn = sinf(-ph);
  v3 rz0 = { c, -sn, 0.0f }, rz1 = { sn, c, 0.0f }, rz2 = { 0.0f, 0.0f, 1.0f };
  v3 t = { dot3(rz0, i), dot3(rz1, i), dot3(rz2, i) };
  float cy = cosf(-th), sy = sinf(-th);
  v3 ry0 = { cy, 0.0f, sy }, ry1 = { 0.0f, 1.0f, 0.0f }, ry2 = { -sy, 0.0f, cy };
  v3 d = { dot3(ry0, t), dot3(ry1, t), dot3(ry2, t) };
  float pd = phi3(d), td = theta3(d);
  if(dot3(i, o) > (float)(1.0 - (double)EPS_F)) pd = 0.0f;
  if(pd >= PI_F) pd = pd - PI_F;
  float iDp = floorf((pd / PI_F + EPS_F) * 180.0f);
  float iDt = floorf((td / HALFPI_F + EPS_F) * 90.0f);
  float a = th / HALFPI_F + EPS_F;
  float iHt = floorf(sqrtf(a < 0.0f ? 0.0f : a) * 90.0f);
  iDp = clampf(iDp, 0.0f, 179.0f); iDt = clampf(iDt, 0.0f, 89.0f); iHt = clampf(iHt, 0.0f, 89.0f);
  float idx = (iHt*90.0f + iDt)*180.0f + iDp;
  if(!(idx == idx)) return 0xFFFFFFFFu;
  return (uint32_t)idx;
}

static float fresnel_cook(float eta, float c)
{
  float t = eta*eta + c*c - 1.0f;
  float g = sqrtf(t < 0.0f ? 0.0f : t);
  float a = (g - c) / (g + c);
  float b = (c*(g + c) - 1.0f) / (c*(g - c) + 1.0f);
  return (float)fmax((double)(0.5f * (a*a) * (1.0f + b*b)), 0.0);
}
static float tan_theta2(v3 v) { float s2 = 1.0f - v.z*v.z; if(s2 < 0.0f) s2 = 0.0f; return s2 / (v.z*v.z); }

/* model: 0 Lambertian(albedo[3]); 1 CookTorrance(albedo[3], roughness, eta); 2 GGX(albedo[3], roughness, eta).
 * component: bsdf_flag bits (1 diffuse, 2 specular) */
void bbmo_eval(int model, const float* a, int component, const float* in, const float* out, float* rgb)
{
  v3 i = { in[0], in[1], in[2] }, o = { out[0], out[1], out[2] };
  rgb[0] = rgb[1] = rgb[2] = 0.0f;
  if(model == 0)
  {
    if(!(component & 1) || !(i.z >= 0.0f && o.z >= 0.0f)) return;
    for(int k = 0; k < 3; ++k) rgb[k] = a[k] / PI_F;
    return;
  }
  if(!(component & 2) || !(i.z > 0.0f && o.z > 0.0f)) return;
  v3 s = { i.x + o.x, i.y + o.y, i.z + o.z };
  v3 h = normalize3(s);
  float ih = dot3(i, h), oh = dot3(o, h);
  float al = a[3], eta = a[4];
  float D = 0.0f, G = 0.0f, nrm;
  if(model == 1)
  {
    if(h.z > 0.0f) { float c2 = h.z*h.z, sx = h.x/al, sy = h.y/al; D = expf(-(sx*sx + sy*sy) / c2) / (al*al*c2*c2); }
    if(ih > 0.0f && oh > 0.0f)
    {
      double gi = 2.0*(double)h.z*(double)i.z/(double)ih, go = 2.0*(double)h.z*(double)o.z/(double)oh;
      G = (float)fmin(1.0, fmin(gi, go));
    }
    nrm = 0.0f;                                              /* Cook: N = pi as double, below */
  }
  else
  {
    if(h.z > 0.0f) { float sx = h.x/al, sy = h.y/al, t = (sx*sx + sy*sy) + h.z*h.z; D = 1.0f / (PI_F * (al*al) * (t*t)); }
    if(ih > 0.0f && oh > 0.0f)
    {
      float g1i = (float)(2.0 / (double)(float)(1.0 + sqrt(1.0 + (double)((al*al) * tan_theta2(i)))));
      float g1o = (float)(2.0 / (double)(float)(1.0 + sqrt(1.0 + (double)((al*al) * tan_theta2(o)))));
      G = g1i * g1o;
    }
    nrm = 4.0f;
  }
  float F = fresnel_cook(eta, 0.5f*(ih + oh));
  float dgf = (D * G) * F;
  double N = (model == 1) ? 3.14159265358979323846 : (double)nrm;
  float u = (float)((double)dgf / N / (double)(i.z*o.z));
  for(int k = 0; k < 3; ++k) rgb[k] = u * a[k];
}

/* nganL2 per-sample error: hsum(pow((v - r) * max(cos_i, 0), 2.0)) * sin_i * sin_o */
float bbmo_ngan_l2_term(const float* in, const float* out, const float* v, const float* r)
{
  float c = in[2] > 0.0f ? in[2] : 0.0f;
  double s = 0.0;
  for(int k = 0; k < 3; ++k) { float t = (v[k] - r[k]) * c; s = s + (double)t*(double)t; }
  float si2 = 1.0f - in[2]*in[2], so2 = 1.0f - out[2]*out[2];
  float si = sqrtf(si2 < 0.0f ? 0.0f : si2), so = sqrtf(so2 < 0.0f ? 0.0f : so2);
  return (float)(s * (double)si * (double)so);
}

/* array forms (AoS, n elements) */
void bbmo_merl_index_n(size_t n, const float* in, const float* out, uint32_t* idx) { for(size_t k = 0; k < n; ++k) idx[k] = bbmo_merl_index(in + 3*k, out + 3*k); }
void bbmo_eval_n(int model, const float* a, int component, size_t n, const float* in, const float* out, float* rgb) { for(size_t k = 0; k < n; ++k) bbmo_eval(model, a, component, in + 3*k, out + 3*k, rgb + 3*k); }
