// Adl/AdlConfig.h -- backend switches (reference: Adl/AdlConfig.h:5-13 enables ADL_ENABLE_CL).
// There is exactly one device backend here, CUDA on sm_100a through libb200rs.so; the reference's
// TYPE_CL enumerator is kept as its name so caller code compiles unchanged.
#pragma once
#define ADL_ENABLE_CUDA
