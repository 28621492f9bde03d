// See bbmcu_host.hpp.  Plain C++17, no CUDA: unit-tested on CPU through the C ABI.
#include "bbmcu_host.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

namespace bbmcu_host {

namespace {

constexpr double kEpsF = FLT_EPSILON, kMaxF = FLT_MAX, kMinF = FLT_MIN;
using VD = std::vector<double>;

AttrInfo attr(const std::string& name, int width, int flag, VD def, VD lo, VD hi, int rows = 1)
{
  auto bc = [&](VD v) { if((int)v.size() == 1) v = VD(width, v[0]); return v; };
  return AttrInfo{name, width, rows, flag, bc(def), bc(lo), bc(hi)};
}
// the attribute families of include/bbm/bsdf_attribute.h:57-94
AttrInfo scale(const std::string& n, int flag) { return attr(n, 3, flag, {0.5}, {0.0}, {1.0}); }
AttrInfo rough(const std::string& n, int w, int flag = ATTR_SPECULAR_PARAM) { return attr(n, w, flag, {0.1}, {kEpsF}, {1.0}); }
AttrInfo sharp(const std::string& n, int w) { return attr(n, w, ATTR_SPECULAR_PARAM, {32.0}, {0.0}, {kMaxF}); }
AttrInfo param(const std::string& n, int w, double def = 1.0, double hi = kMaxF, double lo = 0.0, int flag = ATTR_SPECULAR_PARAM) { return attr(n, w, flag, {def}, {lo}, {hi}); }
AttrInfo ior(const std::string& n) { return attr(n, 1, ATTR_SPECULAR_PARAM, {1.3}, {1.0}, {5.0}); }
AttrInfo refl(const std::string& n, int w) { return attr(n, w, ATTR_SPECULAR_PARAM, {0.1}, {0.0}, {1.0}); }
AttrInfo cior(const std::string& n, int w)   // complex ior: [n..., k...]
{
  VD d, l, h;
  for(int i=0; i < w; ++i) { d.push_back(1.3); l.push_back(0.1); h.push_back(5.0); }
  for(int i=0; i < w; ++i) { d.push_back(0.0); l.push_back(0.0); h.push_back(10.0); }
  return AttrInfo{n, 2*w, 2, ATTR_SPECULAR_PARAM, d, l, h};
}

std::vector<ModelInfo> build_table()
{
  const int SS = ATTR_SPECULAR_SCALE, DS = ATTR_DIFFUSE_SCALE, DEP = ATTR_DEPENDENT;
  std::vector<ModelInfo> t;
  auto add = [&](const std::string& name, std::vector<AttrInfo> a) { t.push_back(ModelInfo{(int)t.size(), name, std::move(a)}); };
  auto laf = [&](int w) { return attr("Cxy", w, ATTR_SPECULAR_PARAM, {-0.57735026919}, {kMinF}, {kMaxF}); };
  AttrInfo cz = attr("Cz", 1, ATTR_SPECULAR_PARAM, {0.57735026919}, {kMinF}, {kMaxF});
  // order = ModelId in bbmcu_models.cuh = SURVEY.md section 8(a3)
  add("Lambertian", {scale("albedo", DS)});
  add("OrenNayar", {scale("albedo", DS), rough("roughness", 1, ATTR_DIFFUSE_PARAM)});
  add("Phong", {scale("albedo", SS), sharp("sharpness", 1)});
  add("NganBlinnPhong", {scale("albedo", SS), sharp("sharpness", 1)});
  add("Lafortune", {scale("albedo", SS), laf(2), cz, sharp("sharpness", 1)});
  add("NganLafortune", {scale("albedo", SS), laf(1), cz, sharp("sharpness", 1)});
  add("Ward", {scale("albedo", SS), rough("roughness", 2)});
  add("WardDuer", {scale("albedo", SS), rough("roughness", 2)});
  add("WardDuerGeislerMoroder", {scale("albedo", SS), rough("roughness", 2)});
  add("NganWard", {scale("albedo", SS), rough("roughness", 1)});
  add("NganWardDuer", {scale("albedo", SS), rough("roughness", 1)});
  add("AshikhminShirley", {refl("fresnelReflectance", 3), sharp("sharpness", 2)});
  add("AshikhminShirleyFull", {scale("diffuseReflectance", DS), refl("fresnelReflectance", 3), sharp("sharpness", 2)});
  add("NganAshikhminShirley", {scale("albedo", SS), refl("fresnelReflectance", 1), sharp("sharpness", 1)});
  add("LowAshikhminShirley", {scale("albedo", SS), ior("fresnelReflectance"), sharp("sharpness", 1)});
  add("CookTorrance", {scale("albedo", SS), rough("roughness", 1), ior("eta")});
  add("LowCookTorrance", {scale("albedo", SS), rough("roughness", 1), ior("eta")});
  add("NganCookTorrance", {scale("albedo", SS), rough("roughness", 1), refl("eta", 1)});
  add("CookTorranceWalter", {scale("albedo", SS), rough("roughness", 1), ior("eta")});
  add("CookTorranceHeitz", {scale("albedo", SS), rough("roughness", 2), ior("eta")});
  add("GGX", {scale("albedo", SS), rough("roughness", 1), ior("eta")});
  add("GGXHeitz", {scale("albedo", SS), rough("roughness", 2), ior("eta")});
  add("PhongWalter", {scale("albedo", SS), sharp("sharpness", 1), ior("eta")});
  add("LowMicrofacet", {param("A", 3), param("B", 1), param("C", 1), ior("eta")});
  add("LowMicrofacetFit", {param("A", 3), param("B", 1), param("C", 1), ior("eta")});
  add("LowSmooth", {param("A", 3), param("B", 1), param("C", 1), ior("eta")});
  add("Ribardiere", {scale("albedo", SS), rough("roughness", 1), param("gamma", 1, 2.0, 40.0, (double)(1.5f + FLT_EPSILON)), ior("eta")});
  add("RibardiereAnisotropic", {scale("albedo", SS), rough("roughness", 2), param("gamma", 1, 2.0, 40.0, (double)(1.5f + FLT_EPSILON)), ior("eta")});
  {
    AttrInfo eta{"eta", 6, 2, ATTR_SPECULAR_PARAM, {1, 1, 1, 0, 0, 0}, {0, 0, 0, -1, -1, -1}, {1, 1, 1, 1, 1, 1}};
    add("Bagher", {scale("albedo", SS), param("K", 3, 7.5, kMaxF, 0.0, DEP), param("Lambda", 3, 1.0, kMaxF, 0.0, DEP),
                   param("c", 3, 1.0, kMaxF, 0.0, DEP), param("theta0", 3, (double)(float)(0.5*M_PI), kMaxF, 0.0, DEP),
                   param("k", 3, 1.0, kMaxF, 0.0, DEP), rough("alpha", 3), param("p", 3, 0.64), eta});
  }
  add("EPD", {param("beta", 1, 0.003, 0.5, 0.0), param("p", 1, 0.2, 5.0, 0.0), cior("eta", 1)});
  add("He", {param("roughness", 1, 0.18), param("autocorrelation", 1, 3.0), cior("eta", 3)});
  add("HeWestin", {param("roughness", 1, 0.18), param("autocorrelation", 1, 3.0), cior("eta", 3)});
  add("HeHolzschuch", {param("roughness", 1, 0.18), param("autocorrelation", 1, 3.0), cior("eta", 3)});
  add("NganHe", {scale("albedo", SS), param("roughness", 1, 0.18), param("autocorrelation", 1, 3.0), ior("eta")});
  // the measured model (include/staticmodel/merl.h): one string attribute (the file name), nothing to fit
  add("Merl", {});
  return t;
}

// ---- string helpers (semantics of include/util/string_util.h) ---------------------------------
std::string trim(const std::string& s)
{
  const char* ws = " \r\n\t\v";
  size_t b = s.find_first_not_of(ws), e = s.find_last_not_of(ws);
  if(b == std::string::npos || b > e) return std::string();
  return s.substr(b, e - b + 1);
}
std::string remove_brackets(const std::string& str)
{
  std::string s = trim(str);
  const std::string open = "[{(", close = "]})";
  bool ok = !s.empty() && open.find(s.front()) != std::string::npos && open.find(s.front()) == close.find(s.back());
  if(!ok) throw std::runtime_error("Mismatch brackets in expression: " + str);
  return s.substr(1, s.size() - 2);
}
std::pair<std::string, std::string> get_keyword(const std::string& str)
{
  size_t b = str.find_first_of('(');
  if(b == std::string::npos) throw std::runtime_error("Expected open bracket in expression: " + str);
  return {b != 0 ? trim(str.substr(0, b)) : std::string(), trim(str.substr(b))};
}
std::pair<std::string, std::string> split_eq(const std::string& str)
{
  size_t p = str.find_first_of('=');
  if(p == std::string::npos) return {std::string(), trim(str)};
  return {trim(str.substr(0, p)), trim(str.substr(p + 1))};
}
std::vector<std::string> split_args(const std::string& str)
{
  const std::string open = "[{(", close = "]})";
  std::vector<size_t> stack;
  std::string word;
  std::vector<std::string> out;
  for(char c : str)
  {
    size_t o = open.find(c), cl = close.find(c);
    if(o != std::string::npos) stack.push_back(o);
    if(cl != std::string::npos)
    {
      if(stack.empty() || stack.back() != cl) throw std::runtime_error("Mismatched brackets in expression: " + str);
      stack.pop_back();
    }
    if(c == ',' && stack.empty()) { out.push_back(trim(word)); word.clear(); }
    else word += c;
  }
  if(!stack.empty()) throw std::runtime_error("Mismatched brackets in expression: " + str);
  if(!trim(word).empty() || !out.empty()) out.push_back(trim(word));
  return out;
}

// which of the reference's two native configurations the host side mirrors while parsing / printing (thread local; set by
// parse_bsdf / import_fit for the duration of a call): floatRGB rounds every value to float, doubleRGB keeps doubles
static thread_local bool g_double_config = false;
struct ConfigScope { bool saved; explicit ConfigScope(bool dbl) : saved(g_double_config) { g_double_config = dbl; } ~ConfigScope() { g_double_config = saved; } };

double parse_scalar(const std::string& s)
{
  size_t pos = 0;
  // doubleRGB: std::stod (core/stringconvert.h:144 is generic over Value); e.g. fits/bagher_sgd.fit needs it (c = 1.3e49)
  if(g_double_config) return std::stod(s, &pos);
  // floatRGB semantics: std::stof (throws std::out_of_range on overflow, e.g. fits/bagher_sgd.fit)
  float v = std::stof(s, &pos);
  return (double)v;
}

// one attribute value: scalar (broadcast), [a, b, c], or [[...], [...]] for two-row attributes
std::vector<double> parse_attr(const AttrInfo& a, const std::string& str)
{
  std::string s = trim(str);
  auto parse_row = [&](const std::string& r, int n) -> std::vector<double> {
    std::vector<double> v;
    std::string t = trim(r);
    if(!t.empty() && t.front() == '[') for(auto& e : split_args(remove_brackets(t))) v.push_back(parse_scalar(e));
    else v.push_back(parse_scalar(t));
    if((int)v.size() == n) return v;
    if(v.size() == 1) return std::vector<double>(n, v[0]);
    throw std::invalid_argument("BBM: too few arguments to convert to std::array<float, " + std::to_string(n) + ">. Found " + std::to_string(v.size()) + " argument in: " + r);
  };
  if(a.rows == 1) return parse_row(s, a.width);
  // two rows of width/2 (complex: [real..., imag...]; Bagher: [F0..., F1...])
  int w = a.width / 2;
  std::vector<std::string> rows;
  if(!s.empty() && s.front() == '[') rows = split_args(remove_brackets(s));
  else rows = {s};
  std::vector<double> out;
  if(rows.size() == 2) { for(auto& r : rows) { auto v = parse_row(r, w); out.insert(out.end(), v.begin(), v.end()); } }
  else if(rows.size() == 1)
  {
    // scalar-to-complex conversion (real part only) is what ior -> complex does (core/ior.h:74-78)
    auto v = parse_row(rows[0], w);
    out = v; out.insert(out.end(), w, 0.0);
  }
  else throw std::invalid_argument("BBM: cannot convert attribute " + a.name + " from: " + str);
  return out;
}

std::string format_attr(const AttrInfo& a, const double* v)
{
  auto row = [&](const double* p, int n) {
    if(n == 1) return format_float(p[0]);
    std::string r = "[";
    for(int i=0; i < n; ++i) { if(i) r += ", "; r += format_float(p[i]); }
    return r + "]";
  };
  if(a.rows == 1) return row(v, a.width);
  int w = a.width / 2;
  return "[" + row(v, w) + ", " + row(v + w, w) + "]";
}

Lobe parse_lobe(const std::string& str)
{
  auto kw = get_keyword(str);
  const ModelInfo* m = find_model(kw.first);
  if(!m) throw std::invalid_argument("BBM: unrecognized BSDF model: " + kw.first + " in: " + str);
  Lobe l; l.model = m;
  if(m->name == "Merl")
  {
    // Merl("file") / Merl(filename = "file"): string_converter<std::string>::fromString strips quotes and blanks
    // (include/core/stringconvert.h:233-245)
    std::string arg = remove_brackets(kw.second);
    auto kv = split_eq(arg);
    if(!kv.first.empty() && kv.first != "filename") throw std::invalid_argument("BBM: invalid argument name: " + kv.first + "(" + kv.second + ") in: " + kw.second);
    const char* quotes = "\"' \r\n\t\v";
    size_t b = kv.second.find_first_not_of(quotes), e = kv.second.find_last_not_of(quotes);
    if(b == std::string::npos || b > e) throw std::invalid_argument("BBM: connect convert to string; unbalanced quotes in: " + kv.second);
    auto data = std::make_shared<MerlData>();
    data->filename = kv.second.substr(b, e - b + 1);
    data->rgb = read_merl(data->filename);
    l.merl = data;
    return l;
  }
  auto args = split_args(remove_brackets(kw.second));
  const size_t NA = m->attrs.size();
  if(args.size() > NA) throw std::invalid_argument("BBM: expected at most " + std::to_string(NA) + " arguments, found " + std::to_string(args.size()) + " in: " + kw.second);
  std::map<std::string, size_t> name_map;
  std::vector<bool> named;
  std::vector<std::string> values;
  for(size_t i=0; i < args.size(); ++i)
  {
    auto kv = split_eq(args[i]);
    if(!kv.first.empty())
    {
      named.push_back(true);
      bool known = false;
      for(auto& a : m->attrs) known |= (a.name == kv.first);
      if(!known) throw std::invalid_argument("BBM: invalid argument name: " + kv.first + "(" + kv.second + ") in: " + kw.second);
      name_map[kv.first] = i;
    }
    else named.push_back(false);
    values.push_back(kv.second);
  }
  for(size_t i=0; i < NA; ++i)
  {
    const AttrInfo& a = m->attrs[i];
    auto it = name_map.find(a.name);
    std::vector<double> v;
    if(it != name_map.end()) v = parse_attr(a, values[it->second]);
    else if(i < values.size() && !named[i]) v = parse_attr(a, values[i]);
    else { v.resize(a.width); for(int k=0; k < a.width; ++k) v[k] = (double)(float)a.def[k]; }
    l.values.insert(l.values.end(), v.begin(), v.end());
  }
  return l;
}

} // anonymous namespace

const std::vector<ModelInfo>& model_table() { static const std::vector<ModelInfo> t = build_table(); return t; }

const ModelInfo* find_model(const std::string& name)
{
  for(auto& m : model_table()) if(m.name == name) return &m;
  return nullptr;
}

std::string format_float(double v)
{
  std::stringstream ss;
  if(g_double_config) ss << v; else ss << (float)v;
  return ss.str();
}

Bsdf parse_bsdf(const std::string& str, bool double_config)
{
  ConfigScope scope(double_config);
  Bsdf b;
  b.double_config = double_config;
  auto kw = get_keyword(trim(str));
  if(kw.first == "Aggregate")
  {
    b.aggregate = true;
    for(auto& a : split_args(remove_brackets(kw.second))) b.lobes.push_back(parse_lobe(a));
    if(b.lobes.empty()) throw std::invalid_argument("BBM: empty Aggregate in: " + str);
  }
  else b.lobes.push_back(parse_lobe(trim(str)));
  return b;
}

std::string Bsdf::to_string() const
{
  ConfigScope scope(double_config);
  auto one = [](const Lobe& l) {
    if(l.merl) return l.model->name + "(\"" + l.merl->filename + "\")";          // merl_data::toString (merl.h:161-164)
    std::string s = l.model->name + "(";
    int off = 0;
    for(size_t i=0; i < l.model->attrs.size(); ++i)
    {
      const AttrInfo& a = l.model->attrs[i];
      if(i) s += ", ";
      s += a.name + " = " + format_attr(a, l.values.data() + off);
      off += a.width;
    }
    return s + ")";
  };
  if(!aggregate) return one(lobes[0]);
  std::string s = "Aggregate(";
  for(size_t i=0; i < lobes.size(); ++i) { if(i) s += ", "; s += one(lobes[i]); }
  return s + ")";
}

int Bsdf::param_count(int flags) const
{
  int n = 0;
  for(auto& l : lobes) for(auto& a : l.model->attrs) if(a.flag & flags) n += a.width;
  return n;
}

std::vector<double> Bsdf::params(int which, int flags) const
{
  std::vector<double> out;
  for(auto& l : lobes)
  {
    int off = 0;
    for(auto& a : l.model->attrs)
    {
      if(a.flag & flags)
        for(int k=0; k < a.width; ++k)
          out.push_back(which == 0 ? l.values[off + k] : which == 1 ? (double)(float)a.def[k] : which == 2 ? (double)(float)a.lo[k] : (double)(float)a.hi[k]);
      off += a.width;
    }
  }
  return out;
}

void Bsdf::set_params(int flags, const double* v, int n)
{
  if(n != param_count(flags)) throw std::invalid_argument("BBM: parameter count mismatch: expected " + std::to_string(param_count(flags)) + ", got " + std::to_string(n));
  int j = 0;
  for(auto& l : lobes)
  {
    int off = 0;
    for(auto& a : l.model->attrs)
    {
      if(a.flag & flags) for(int k=0; k < a.width; ++k) l.values[off + k] = v[j++];
      off += a.width;
    }
  }
}

// merl_data::import (include/staticmodel/merl.h:173-206): three u32 dimensions (90, 90, 180), then three planes of
// doubles; negative values clamp to 0; channels scale by 1/1500, 1.15/1500, 1.66/1500; stored as double, read as float
std::vector<float> read_merl(const std::string& filename)
{
  std::ifstream ifs(filename.c_str(), std::ios_base::binary);
  if(!ifs) throw std::runtime_error("BBM: unable to open MERL BRDF: \"" + filename + "\"");
  uint32_t dims[3] = {0, 0, 0};
  ifs.read(reinterpret_cast<char*>(dims), sizeof(dims));
  if(!ifs || dims[0] != 90 || dims[1] != 90 || dims[2] != 180) throw std::runtime_error("BBM: not a recognized MERL BRDF: \"" + filename + "\"");
  const size_t N = size_t(90)*90*180;
  std::vector<double> buf(3*N);
  ifs.read(reinterpret_cast<char*>(buf.data()), 3*N*sizeof(double));
  if(!ifs) throw std::runtime_error("BBM: truncated MERL BRDF: \"" + filename + "\"");
  const double scale[3] = {1.0, 1.15, 1.66};
  std::vector<float> rgb(3*N);
  for(int c=0; c < 3; ++c)
    for(size_t i=0; i < N; ++i)
      rgb[c*N + i] = (float)std::fmax(0.0, buf[c*N + i] * scale[c] / 1500.0);
  return rgb;
}

void write_merl(const std::string& filename, const float* rgb)
{
  std::ofstream ofs(filename.c_str(), std::ios_base::binary);
  if(!ofs) throw std::runtime_error("BBM: unable to write MERL BRDF: " + filename);
  uint32_t dims[3] = {90, 90, 180};
  ofs.write(reinterpret_cast<const char*>(dims), sizeof(dims));
  const size_t N = size_t(90)*90*180;
  std::vector<double> buf(3*N);
  const double scale[3] = {1.0, 1.15, 1.66};
  for(int c=0; c < 3; ++c) for(size_t i=0; i < N; ++i) buf[c*N + i] = (double)rgb[c*N + i] * 1500.0 / scale[c];
  ofs.write(reinterpret_cast<const char*>(buf.data()), 3*N*sizeof(double));
}

std::vector<std::pair<std::string, Bsdf>> import_fit(const std::string& filename, bool double_config)
{
  std::ifstream ifs(filename.c_str());
  if(!ifs) throw std::runtime_error("BBM: unable to open FIT file: " + filename);
  // the reference stores into a std::map: sorted by key, first occurrence wins (io/fit.h:44-52)
  std::map<std::string, Bsdf> m;
  for(std::string line; std::getline(ifs, line);)
  {
    size_t pos = line.find('#');
    if(pos == 0 || line.empty()) continue;
    if(pos != std::string::npos) line = line.substr(0, pos);
    auto kv = split_eq(line);
    if(kv.first.empty()) continue;
    try { m.emplace(kv.first, parse_bsdf(kv.second, double_config)); }
    catch(const std::exception& e) { throw std::invalid_argument(std::string(e.what()) + " (key " + kv.first + ")"); }
  }
  return std::vector<std::pair<std::string, Bsdf>>(m.begin(), m.end());
}

void export_fit(const std::string& filename, const std::vector<std::pair<std::string, Bsdf>>& data, const std::string& comment)
{
  std::ofstream ofs(filename.c_str());
  if(!ofs) throw std::runtime_error("BBM: unable to write FIT file: " + filename);
  std::stringstream ss(comment);
  for(std::string line; std::getline(ss, line);) ofs << "# " << line << std::endl;
  std::map<std::string, const Bsdf*> m;
  for(auto& d : data) m.emplace(d.first, &d.second);
  for(auto& d : m) ofs << d.first << " = " << d.second->to_string() << std::endl;
}

} // namespace bbmcu_host
