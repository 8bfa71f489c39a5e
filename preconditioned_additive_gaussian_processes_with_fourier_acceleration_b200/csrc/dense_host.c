/* Dense additive kernel handles (placeholder; see precond_setup.c). */
#include <stdio.h>
#include <stdlib.h>
#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

void *Nfft4GPKernelAdditiveKernelParamCreate(double *data, int n, int ldim, int d, int *windows, int nwindows, int dwindows, func_kernel fkernel)
{
   (void)data; (void)n; (void)ldim; (void)d; (void)windows; (void)nwindows; (void)dwindows; (void)fkernel;
   return NULL;
}
int Nfft4GPKernelAdditiveKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   (void)str; (void)data; (void)n; (void)ldim; (void)d; (void)permr; (void)kr; (void)permc; (void)kc; (void)Kp; (void)dKp;
   return -1;
}
int Nfft4GPKernelGaussianKernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   (void)str; (void)data; (void)n; (void)ldim; (void)d; (void)permr; (void)kr; (void)permc; (void)kc; (void)Kp; (void)dKp;
   return -1;
}
int Nfft4GPKernelMatern12Kernel(void *str, double *data, int n, int ldim, int d, int *permr, int kr, int *permc, int kc, double **Kp, double **dKp)
{
   (void)str; (void)data; (void)n; (void)ldim; (void)d; (void)permr; (void)kr; (void)permc; (void)kc; (void)Kp; (void)dKp;
   return -1;
}
void Nfft4GPKernelParamFree(void *str) { free(str); }
