#ifndef _BBM_CUDA_LOSS_H_
#define _BBM_CUDA_LOSS_H_

/************************************************************************/
/*! \file bbm_cuda/loss.h

  \brief The six fitting metrics of include/loss/ on the CUDA backbone.

  bbm::cuda::loss<FITTED> stands where the reference's loss classes stand
  (nganL2, lowL2, bieronL2, lowLog, bieronLog, standardLog; include/loss/
  cosine_weighted_l2.h:52-78, cosine_weighted_log.h:61-87): it is constructed
  from the LIVE fitted model and a reference, holds the fitted model by
  reference (include/bbm/sampledlossfunction.h:93-94), and satisfies
  concepts::lossfunction and concepts::sampledlossfunction - so the UNMODIFIED
  bbm::compass (include/optimizer/compass.h:39-185) drives it: compass writes
  the model's parameters through parameter_values(fitted), calls update() and
  operator()(mask); every operator() reads the live parameters and runs ONE
  kernel pass over all samples of the linearizer on the GPU.

  The model reaches the device as its own toString() (the BSDF string grammar,
  include/core/stringconvert.h:512-576) once, at construction; afterwards only
  the parameter vector travels.

  Needs the reference's headers and include/bbmcu/loss.hpp of this repository.
*************************************************************************/

#include <string>
#include <vector>
#include <optional>

#include "concepts/sampledlossfunction.h"
#include "bbm/config.h"
#include "bbm/bsdf_enumerate.h"
#include "core/stringconvert.h"

#include "bbmcu/loss.hpp"

namespace bbm {
  namespace cuda {

    //! \brief the metrics, named as the reference's loss classes
    using metric = ::bbmcu::metric;

    /********************************************************************/
    /*! \brief A sampled loss function evaluated on the GPU

      \tparam FITTED = the bbm BSDF model type being fitted (any concepts::bsdfmodel
                       whose toString() the CUDA backbone can import)

      Satisfies: concepts::sampledlossfunction
    *********************************************************************/
    template<typename FITTED> requires concepts::has_config<FITTED>
      class loss
    {
    public:
      BBM_IMPORT_CONFIG( FITTED );

      /******************************************************************/
      /*! \brief Loss over spherical_linearizer(samplesIn, samplesOut) against an analytic reference

        \param ctx = CUDA context (device + stream)
        \param m = which of the six metrics
        \param fitted = the live model (kept by reference)
        \param reference = reference BSDF (anything with a toString the CUDA backbone can import)
        \param samplesIn, samplesOut = (phi, theta) sample counts, as in nganL2(fitted, reference, samplesIn, samplesOut)
      *******************************************************************/
      template<typename REFERENCE>
        loss(const ::bbmcu::context& ctx, metric m, FITTED& fitted, const REFERENCE& reference, const vec2d<Size_t>& samplesIn, const vec2d<Size_t>& samplesOut)
        : _ctx(ctx), _fitted(fitted), _live(bbm::parameter_values(fitted)), _shape(ctx, bbm::toString(fitted)), _reference(ctx, bbm::toString(reference)), _params(std::size(_live))
      {
        _loss.emplace(m, _shape, _reference, _params, ::bbmcu::spherical_grid({uint32_t(samplesIn[0]), uint32_t(samplesIn[1])}, {uint32_t(samplesOut[0]), uint32_t(samplesOut[1])}));
      }

      //! \brief Loss over the MERL linearizer (90 x 90 x 180) against a measured table (3 planes of 1 458 000 floats: merl<> data, scaled as merl.h:173-206 scales it)
      loss(const ::bbmcu::context& ctx, metric m, FITTED& fitted, const float* merl_rgb_planes)
        : _ctx(ctx), _fitted(fitted), _live(bbm::parameter_values(fitted)), _shape(ctx, bbm::toString(fitted)), _params(std::size(_live))
      {
        _loss.emplace(m, _shape, merl_rgb_planes, _params);
      }

      //! \brief the reference tabulation is fixed at construction; the parameters are read at every evaluation
      void update(void) {}

      //! \brief loss of the live model (one GPU pass over all samples); 0 when masked (sampledlossfunction.h:65-66)
      Value operator()(Mask mask=true) const
      {
        if(!mask) return 0;
        pull();
        return Value( (*_loss)(true) );
      }

      //! \brief loss of sample 'idx' (sampledlossfunction.h:62-73): terms of all samples are computed in one pass and cached until the parameters change
      Value operator()(const Size_t& idx, Mask mask=true) const
      {
        if(!mask) return 0;
        pull();
        if(_terms.empty() || _terms_of != _params)
        {
          _terms.resize(samples());
          ::bbmcu::check(bbmcu_bsdf_set_params(_shape.get(), BBMCU_ATTR_ALL, _params.data(), int(_params.size())));
          ::bbmcu::check(bbmcu_loss_terms(_loss->handle(), _shape.get(), _terms.data()), _ctx.get());
          _terms_of = _params;
        }
        return Value(_terms.at(idx));
      }

      //! \brief number of samples of the linearizer
      Size_t samples(void) const { return Size_t(_loss->samples()); }

      /******************************************************************/
      /*! @{ \name New on this backbone (the reference has no gradient, README.md:27-28)
       ******************************************************************/
      //! \brief loss and d loss / d parameter (order of parameter_values(fitted)) of the live model
      Value gradient(std::vector<Value>& grad) const
      {
        pull();
        std::vector<double> g;
        double l = _loss->gradient(g);
        grad.assign(g.begin(), g.end());
        return Value(l);
      }
      //! \brief losses of K parameter vectors (row-major K x P) in ONE launch - e.g. all 2P probes of a compass step
      std::vector<double> batch(const std::vector<double>& params_KxP) const { return (*_loss)(params_KxP); }
      //! \brief the loss object of include/bbmcu/loss.hpp underneath (peer exchange, device results)
      ::bbmcu::cuda_loss& backend(void) { return *_loss; }
      //! @}

    private:
      //! \brief copy the live parameters (references into the model) into the vector the backend reads
      void pull(void) const { size_t j=0; for(const auto& p : _live) _params[j++] = double(Value(p)); }

      ::bbmcu::context _ctx;
      FITTED& _fitted;
      decltype(bbm::parameter_values(std::declval<FITTED&>())) _live;
      mutable ::bbmcu::cuda_bsdf _shape;
      ::bbmcu::cuda_bsdf _reference;
      mutable std::vector<double> _params;
      mutable std::optional<::bbmcu::cuda_loss> _loss;
      mutable std::vector<float> _terms;
      mutable std::vector<double> _terms_of;
    };

    // (concepts::lossfunction / sampledlossfunction / optimization_algorithm<compass<loss<...>>> are asserted for a concrete
    //  model in tests/cpp/test_reference_boundary.cpp: the class template needs a model with reflected attributes)

  } // end cuda namespace
} // end bbm namespace

#endif /* _BBM_CUDA_LOSS_H_ */
