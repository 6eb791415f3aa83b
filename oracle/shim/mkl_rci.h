// TEST INFRASTRUCTURE: forwards to the MKL stand-in (see mkl_shim.h)
#pragma once
#include "mkl_shim.h"
