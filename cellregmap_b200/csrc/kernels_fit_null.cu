// Translation unit: K2 instantiations for the covariates-only design (association null model).
#define CRM_FIT_NULL_TU 1
#include "kernels_fit_g.cu"
