#ifndef _BBM_CUDA_BACKBONE_H_
#define _BBM_CUDA_BACKBONE_H_

/************************************************************************/
/*! \file backbone.h

  \brief The CUDA backbone: configurations.

  Scalar host code runs on the native backbone's types - this header pulls in
  backbone/native/include/backbone.h unchanged (it must come later on the include
  path) - and adds one configuration, floatRGB_cuda, with the same Value and
  Spectrum as floatRGB.  Every bbm template instantiates for it exactly as it does
  for floatRGB, and BBM_VALIDATE_BACKBONE (include/core/backbone.h:34-49) passes
  because the types ARE the native backbone's.  The tag is what the adapters of
  include/bbm_cuda/ key on: objects of this configuration evaluate batches, losses
  and gradients through libbbmcu.so (include/bbmcu.h) instead of the scalar loop.
*************************************************************************/

#include_next "backbone.h"        // backbone/native/include/backbone.h: floatRGB, doubleRGB, detail::rgbConfig

namespace bbm {

  /*** floatRGB on the host, batches on the B200 ***/
  struct floatRGB_cuda : public detail::rgbConfig<float, "floatRGB_cuda", floatRGB_cuda> {};

} // end bbm namespace

#endif /* _BBM_CUDA_BACKBONE_H_ */
