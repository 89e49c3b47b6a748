// forwarding header of the minimal ITK stand-in (tests/itk_shim/itkShim.h)
#include "itkShim.h"
