/* Preconditioner setups (C host side).  See DESIGN.md for the status of each. */
#include <stdio.h>
#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"
#include "b200_precond.h"

int Nfft4GPPrecondNysSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad, void *vnys_mat)
{
   (void)data; (void)n; (void)ldim; (void)d; (void)fkernel; (void)fkernel_params; (void)require_grad; (void)vnys_mat;
   printf("Error: nfft4gp_b200: Nystrom setup not built yet.\n");
   return -1;
}
int Nfft4GPPrecondFsaiSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad, void *vfsai_mat)
{
   (void)data; (void)n; (void)ldim; (void)d; (void)fkernel; (void)fkernel_params; (void)require_grad; (void)vfsai_mat;
   printf("Error: nfft4gp_b200: FSAI setup not built yet.\n");
   return -1;
}
int Nfft4GPPrecondAFNSetupWithKernel(double *data, int n, int ldim, int d, func_kernel fkernel, void *fkernel_params, int require_grad, void *vafn_mat)
{
   (void)data; (void)n; (void)ldim; (void)d; (void)fkernel; (void)fkernel_params; (void)require_grad; (void)vafn_mat;
   printf("Error: nfft4gp_b200: AFN setup not built yet.\n");
   return -1;
}
