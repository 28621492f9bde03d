// bbmcu/optimizer.hpp - optimizers behind the reference's optimization_algorithm interface
// (include/concepts/optimization_algorithm.h:24-37: step() -> loss, reset(), is_converged()).
//
//  * compass          the reference's compass search (include/optimizer/compass.h:82-140; Kolda et al. 2003, p. 402) restated:
//                     same probe order (+1, -1, +2, -2, ...), same box test, first strictly better probe wins ties, same
//                     contraction / expansion rule and the same convergence test (step < tolerance).  It runs against any
//                     concepts::lossfunction, one launch per probe, exactly like the reference.
//  * compass_batched  the same search with all 2P probes of a step evaluated in ONE launch (K = 2P parameter vectors,
//                     one loss pass is ~3 us of work, so launch latency dominates otherwise).  Probes are
//                     formed from the un-probed base vector, so the reference's float drift of "p + s - s" is not reproduced;
//                     trajectories agree until float-sum noise decides a tie (compare final losses).
//  * gradient_descent Adam on the analytic gradient, projected onto the box (new capability; the reference has no gradient).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include "loss.hpp"

namespace bbmcu {

template<class LOSS>
class compass
{
public:
  compass(LOSS& loss, std::vector<double>& param, std::vector<double> lower = {}, std::vector<double> upper = {},
          float tolerance = std::numeric_limits<float>::epsilon(), float stepSize = 1.0f, float contraction = 0.5f, float expansion = 1.0f)
    : _loss(loss), _param(param), _lower(std::move(lower)), _upper(std::move(upper)), _initialStep(stepSize), _tolerance(tolerance),
      _contraction(contraction), _expansion(expansion)
  { reset(); }

  float step()
  {
    if(is_converged()) return 0.0f;
    _loss.update();
    int best = 0;
    float loss = _lossValue;
    const int P = (int)_param.size();
    for(int i = 1; i <= P; ++i)
      for(int sgn = +1; sgn >= -1; sgn -= 2)
      {
        const int cardinal = sgn * i;
        probe(cardinal);
        if(in_box())
        {
          float err = _loss(true);
          if(err < loss) { best = cardinal; loss = err; }
        }
        probe(-cardinal);
      }
    const bool improved = loss < _lossValue;
    if(improved) { probe(best); _lossValue = loss; }
    _step = improved ? _expansion * _step : _contraction * _step;
    return _lossValue;
  }
  void reset() { _step = _initialStep; _loss.update(); _lossValue = _loss(true); }
  bool is_converged() const { return _step < _tolerance; }
  float loss() const { return _lossValue; }
  float step_size() const { return _step; }

private:
  // parameters are floats in the reference's floatRGB configuration: value + step is a float addition
  void probe(int cardinal) { if(cardinal == 0) return; double& p = _param[std::abs(cardinal) - 1]; p = (double)((float)p + (cardinal < 0 ? -_step : _step)); }
  bool in_box() const
  {
    if(_lower.empty() && _upper.empty()) return true;
    for(size_t j = 0; j < _param.size(); ++j)
    {
      if(!_lower.empty() && !((float)_param[j] >= (float)_lower[j])) return false;
      if(!_upper.empty() && !((float)_param[j] <= (float)_upper[j])) return false;
    }
    return true;
  }
  LOSS& _loss;
  std::vector<double>& _param;
  std::vector<double> _lower, _upper;
  float _initialStep, _tolerance, _contraction, _expansion, _step = 1.0f, _lossValue = 0.0f;
};

class compass_batched
{
public:
  compass_batched(cuda_loss& loss, std::vector<double> lower = {}, std::vector<double> upper = {},
                  float tolerance = std::numeric_limits<float>::epsilon(), float stepSize = 1.0f, float contraction = 0.5f, float expansion = 1.0f)
    : _loss(loss), _param(loss.live_parameters()), _lower(std::move(lower)), _upper(std::move(upper)), _initialStep(stepSize), _tolerance(tolerance),
      _contraction(contraction), _expansion(expansion)
  { reset(); }

  float step()
  {
    if(is_converged()) return 0.0f;
    const size_t P = _param.size();
    _probes.assign(2 * P * P, 0.0);
    // the reference tests ALL parameters of a probe against the box (compass.h:117-121), not only the one it moved: a
    // start point outside the box in another coordinate rejects every probe
    bool base_inside = true;
    for(size_t i = 0; i < P; ++i)
    {
      const float v = (float)_param[i];
      if(!_lower.empty() && !(v >= (float)_lower[i])) base_inside = false;
      if(!_upper.empty() && !(v <= (float)_upper[i])) base_inside = false;
    }
    std::vector<char> ok(2 * P, 1);
    for(size_t k = 0; k < 2 * P; ++k)
    {
      const size_t j = k / 2;
      if(!base_inside)
      {
        // with the other coordinates outside, only a probe whose own coordinate is the single offender can be inside
        bool others = true;
        for(size_t i = 0; i < P; ++i) if(i != j) { const float v = (float)_param[i]; if((!_lower.empty() && !(v >= (float)_lower[i])) || (!_upper.empty() && !(v <= (float)_upper[i]))) others = false; }
        if(!others) ok[k] = 0;
      }
      std::copy(_param.begin(), _param.end(), _probes.begin() + k * P);
      float v = (float)_param[j] + ((k & 1) ? -_step : _step);
      _probes[k * P + j] = (double)v;
      if(!_lower.empty() && !(v >= (float)_lower[j])) ok[k] = 0;
      if(!_upper.empty() && !(v <= (float)_upper[j])) ok[k] = 0;
      if(!ok[k]) _probes[k * P + j] = _param[j];          // evaluated but ignored (keeps the launch shape fixed)
    }
    std::vector<double> err = _loss(_probes);              // ONE launch for all 2P probes
    int best = -1;
    float loss = _lossValue;
    for(size_t k = 0; k < 2 * P; ++k) if(ok[k] && (float)err[k] < loss) { best = (int)k; loss = (float)err[k]; }
    const bool improved = best >= 0;
    if(improved) { _param[best / 2] = _probes[(size_t)best * P + best / 2]; _lossValue = loss; }
    _step = improved ? _expansion * _step : _contraction * _step;
    return _lossValue;
  }
  void reset() { _step = _initialStep; _lossValue = _loss(true); }
  bool is_converged() const { return _step < _tolerance; }
  float loss() const { return _lossValue; }

private:
  cuda_loss& _loss;
  std::vector<double>& _param;
  std::vector<double> _lower, _upper, _probes;
  float _initialStep, _tolerance, _contraction, _expansion, _step = 1.0f, _lossValue = 0.0f;
};

class gradient_descent
{
public:
  gradient_descent(cuda_loss& loss, std::vector<double> lower = {}, std::vector<double> upper = {}, double rate = 1e-2, double tolerance = 1e-7)
    : _loss(loss), _param(loss.live_parameters()), _lower(std::move(lower)), _upper(std::move(upper)), _rate(rate), _tolerance(tolerance)
  { reset(); }
  float step()
  {
    std::vector<double> g;
    double l = _loss.gradient(g);
    ++_t;
    double moved = 0.0;
    for(size_t j = 0; j < _param.size(); ++j)
    {
      if(!std::isfinite(g[j])) continue;
      _m[j] = 0.9 * _m[j] + 0.1 * g[j];
      _v[j] = 0.999 * _v[j] + 0.001 * g[j] * g[j];
      double mh = _m[j] / (1.0 - std::pow(0.9, _t)), vh = _v[j] / (1.0 - std::pow(0.999, _t));
      double scale = std::max(1e-3, std::abs(_param[j]));                     // relative steps: parameters span 1e-3 .. 1e4
      double p = _param[j] - _rate * scale * mh / (std::sqrt(vh) + 1e-12);
      if(!_lower.empty()) p = std::max(p, _lower[j]);
      if(!_upper.empty()) p = std::min(p, _upper[j]);
      moved = std::max(moved, std::abs(p - _param[j]) / scale);
      _param[j] = p;
    }
    _converged = moved < _tolerance;
    _lossValue = (float)l;
    return _lossValue;                                                         // loss BEFORE this update
  }
  void reset() { _m.assign(_param.size(), 0.0); _v.assign(_param.size(), 0.0); _t = 0; _converged = false; }
  bool is_converged() const { return _converged; }
  float loss() const { return _lossValue; }
private:
  cuda_loss& _loss;
  std::vector<double>& _param;
  std::vector<double> _lower, _upper, _m, _v;
  double _rate, _tolerance;
  int _t = 0;
  bool _converged = false;
  float _lossValue = 0.0f;
};

} // namespace bbmcu
