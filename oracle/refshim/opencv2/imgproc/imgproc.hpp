/* reference-build shim: see ../../minicv.hpp */
#include "../../minicv.hpp"
