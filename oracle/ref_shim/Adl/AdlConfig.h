/*
 * oracle/ref_shim/Adl/AdlConfig.h -- TEST INFRASTRUCTURE.
 * Shadows /root/reference/Adl/AdlConfig.h (which hard-codes ADL_ENABLE_CL at line 5) when the
 * reference's Host backend is compiled as the oracle: this image has no OpenCL headers, so no
 * GPU backend of the reference is enabled.  Put -Ioracle/ref_shim BEFORE -I/root/reference.
 */
#pragma once
/* intentionally no ADL_ENABLE_CL / ADL_ENABLE_DX11 */
