// Host side of the CUDA backbone: the model/attribute table (the reference's compile-time
// reflection flattened into data), the BSDF string grammar, parameter enumeration and .fit I/O.
//
// Contract restated from the reference:
//   attribute order   include/util/reflection.h:141-148 (own attributes first, then bases)
//   enumeration       include/bbm/bsdf_enumerate.h:39-47,102-237, include/core/enumerate.h:28-42
//   defaults/bounds   include/bbm/bsdf_attribute.h:57-94 and each model header
//   string grammar    include/core/stringconvert.h:61-75,333-360,512-576, include/util/string_util.h
//   .fit              include/io/fit.h:34-77
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace bbmcu_host {

// bsdf_attr flags (include/bbm/bsdf_attr_flag.h:17-31)
enum : int { ATTR_NONE = 0, ATTR_DIFFUSE_SCALE = 1, ATTR_DIFFUSE_PARAM = 2, ATTR_SPECULAR_SCALE = 4, ATTR_SPECULAR_PARAM = 8,
             ATTR_DEPENDENT = 16, ATTR_ALL = 15 };

struct AttrInfo
{
  std::string name;
  int width;                 // number of scalars
  int rows;                  // 1, or 2 for [[..],[..]] attributes (complex RGB ior, Bagher F0/F1)
  int flag;                  // bsdf_attr bit
  std::vector<double> def, lo, hi;   // per scalar (double so that doubleRGB round trips survive)
};

struct ModelInfo
{
  int id;
  std::string name;
  std::vector<AttrInfo> attrs;
  int n_floats() const { int n = 0; for(auto& a : attrs) n += a.width; return n; }
};

const std::vector<ModelInfo>& model_table();
const ModelInfo* find_model(const std::string& name);

// a loaded MERL measurement (include/staticmodel/merl.h:173-206): 3 planes of 1 458 000 floats on the host, plus
// lazily created per-device copies (uploaded and released by the CUDA side through the two hooks, so that this file
// stays free of CUDA)
struct MerlData
{
  std::string filename;
  std::vector<float> rgb;
  mutable std::map<int, void*> device;                       // device index -> device copy
  mutable void (*release)(int device, void* ptr) = nullptr;  // set by whoever created the copies
  ~MerlData() { if(release) for(auto& d : device) release(d.first, d.second); }
};
std::vector<float> read_merl(const std::string& filename);                        // throws std::runtime_error like merl_data::import
void write_merl(const std::string& filename, const float* rgb);

struct Lobe
{
  const ModelInfo* model;
  std::vector<double> values;        // full attribute block (all attributes incl. Dependent), reflection order
  std::shared_ptr<const MerlData> merl;   // only for the measured model Merl("file")
};

// bsdf_ptr equivalent: one model, or Aggregate(...) of several
struct Bsdf
{
  bool aggregate = false;
  bool double_config = false;        // parsed / printed as the reference's doubleRGB does (values not rounded to float on the host)
  std::vector<Lobe> lobes;

  std::string to_string() const;
  // forward-order enumeration (the reference's run-time aggregate reverses each
  // lobe; we keep the documented forward order everywhere)
  int param_count(int flags) const;
  std::vector<double> params(int which, int flags) const;     // which: 0 value 1 default 2 lower 3 upper
  void set_params(int flags, const double* v, int n);
  int attr_floats() const { int n = 0; for(auto& l : lobes) n += (int)l.values.size(); return n; }
};

// throws std::invalid_argument / std::runtime_error with the reference's wording where it has one
Bsdf parse_bsdf(const std::string& str, bool double_config = false);
std::string format_float(double v);                            // default ostream << float (6 significant digits)

std::vector<std::pair<std::string, Bsdf>> import_fit(const std::string& filename, bool double_config = false);   // file order preserved
void export_fit(const std::string& filename, const std::vector<std::pair<std::string, Bsdf>>& data, const std::string& comment);

} // namespace bbmcu_host
