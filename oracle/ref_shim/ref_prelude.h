/*
 * oracle/ref_shim/ref_prelude.h -- TEST INFRASTRUCTURE.  Force-included (-include) when compiling
 * the unmodified reference with its CL backend disabled:
 *  - /root/reference/Adl/Host/AdlHost.inl:113,120,139 call memcpy without <string.h>;
 *  - /root/reference/Adl/Adl.inl:49 names adl::DeviceCL unconditionally.
 */
#pragma once
#include <string.h>
namespace adl { struct DeviceCL { int getNCUs() const { return 0; } }; }
