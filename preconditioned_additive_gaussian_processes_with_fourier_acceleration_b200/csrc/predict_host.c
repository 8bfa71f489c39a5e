/* GP prediction through the NFFT additive kernel: the remaining exported entry points of the reference's
 * SRC/external/nfft_interface.c, so that the object can be dropped from the reference build as a whole.
 *   Nfft4GPNFFTAppendData          nfft_interface.c:858-871   stack the training and the test features (feature-major)
 *   Nfft4GPAdditiveNFFTGpPredict   nfft_interface.c:873-1068  mean = K21 K11^{-1} y through ONE solve on the training handle
 *                                                             and ONE matvec on the joint train+test handle; optional
 *                                                             standard deviation: one solve per test point
 * Host-side orchestration in C, statement by statement as the reference (including: precond_setup is called with
 * require_grad = 1 at :936; the mean solve runs FGMRES with kdim = maxits, the per-point solves with kdim = maxits = n;
 * std = sqrt(|K22_ii - k_i' K11^{-1} k_i|)).  Every solve and matvec runs on the GPU through this library's
 * Nfft4GPSolverFgmres and the caller's function pointers.  The hyper-parameter transform (SRC/optimizer/transform.c:4-96)
 * is restated here because the optimizer object is not part of this library. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

/* SRC/optimizer/transform.c:4-96, forward direction only */
static double transform_value(nfft4gp_transform_type type, double val)
{
   switch (type)
   {
      case NFFT4GP_TRANSFORM_SIGMOID:
         return 1.0 / (exp(-val) + 1.0);
      case NFFT4GP_TRANSFORM_SOFTPLUS:
         if (val > 20.0)
            return val;
         if (val < -20.0)
            return exp(val);
         return log(1.0 + exp(val));
      case NFFT4GP_TRANSFORM_EXP:
         return exp(val);
      default:
         return val;
   }
}

double *Nfft4GPNFFTAppendData(double *X1, int n1, int ldim1, int d, double *X2, int n2, int ldim2)
{
   double *X = (double *)b200_xmalloc(sizeof(double) * (size_t)(n1 + n2) * (size_t)d);
   int i;
   for (i = 0; i < d; i++)
   {
      memcpy(X + (size_t)i * (n1 + n2), X1 + (size_t)i * ldim1, sizeof(double) * (size_t)n1);
      memcpy(X + (size_t)i * (n1 + n2) + n1, X2 + (size_t)i * ldim2, sizeof(double) * (size_t)n2);
   }
   return X;
}

int Nfft4GPAdditiveNFFTGpPredict(double *x, double *data, double *label, int n, int ldim, int d, double *data_predict, int n_predict,
                                 int ldim_predict, double *data_all, func_kernel fkernel, void *vfkernel_data, void *vfkernel_data_l,
                                 func_free kernel_data_free, func_symmatvec matvec, func_kernel precond_fkernel, void *precond_vfkernel_data,
                                 func_free precond_kernel_data_free, precond_kernel_setup precond_setup, func_solve precond_solve,
                                 void *precond_data, int atol, double tol, int maxits, nfft4gp_transform_type transform, int print_level,
                                 double *dwork, double **label_predictp, double **std_predictp)
{
   nfft4gp_kernel *kd = (nfft4gp_kernel *)vfkernel_data, *kl = (nfft4gp_kernel *)vfkernel_data_l;
   double tvals[3], *K11_mat = NULL, *dK11_mat = NULL, *K_mat = NULL, *dK_mat = NULL;
   double *label_predict, *label_helper, *iKY, rel_res, *rel_res_v = NULL;
   int i, niter, nall = n + n_predict;
   (void)data_predict; (void)ldim_predict; (void)kernel_data_free; (void)precond_kernel_data_free; (void)dwork;
   for (i = 0; i < 3; i++)
      tvals[i] = transform_value(transform, x[i]);
   kd->_params[0] = kl->_params[0] = tvals[0]; /* f  */
   kd->_params[1] = kl->_params[1] = tvals[1]; /* l  */
   kd->_noise_level = kl->_noise_level = tvals[2]; /* mu */
   if (fkernel(vfkernel_data, data, n, ldim, d, NULL, 0, NULL, 0, &K11_mat, &dK11_mat) ||
       fkernel(vfkernel_data_l, data_all, nall, nall, d, NULL, 0, NULL, 0, &K_mat, &dK_mat))
      return -1;
   if (precond_setup != NULL)
   {
      if (precond_setup(data, n, ldim, d, precond_fkernel, precond_vfkernel_data, 1, precond_data))
         return -1;
   }
   else
      precond_data = NULL;
   label_predict = *label_predictp ? *label_predictp : (double *)b200_xmalloc(sizeof(double) * (size_t)(n_predict ? n_predict : 1));
   iKY = (double *)b200_xcalloc((size_t)nall, sizeof(double));
   label_helper = (double *)b200_xmalloc(sizeof(double) * (size_t)nall);
   if (Nfft4GPSolverFgmres(K11_mat, n, matvec, precond_data, precond_solve, iKY, label, maxits, maxits, atol, tol, &rel_res, &rel_res_v, &niter,
                           print_level))
      return -1;
   /* [K11 K12; K21 K22] [iKY; 0] = [.; K21 iKY] */
   if (matvec(K_mat, nall, 1.0, iKY, 0.0, label_helper))
      return -1;
   memcpy(label_predict, label_helper + n, sizeof(double) * (size_t)n_predict);
   free(rel_res_v);
   rel_res_v = NULL;
   *label_predictp = label_predict;
   if (std_predictp)
   {
      double *std_predict = *std_predictp ? *std_predictp : (double *)b200_xmalloc(sizeof(double) * (size_t)(n_predict ? n_predict : 1));
      for (i = 0; i < n_predict; i++)
      {
         double K22i, dot = 0.0;
         int j;
         /* column n + i of the joint kernel: K12(:, i) in the first n entries, K22(i, i) at n + i */
         memset(iKY, 0, sizeof(double) * (size_t)nall);
         iKY[n + i] = 1.0;
         if (matvec(K_mat, nall, 1.0, iKY, 0.0, label_helper))
            return -1;
         K22i = label_helper[n + i];
         iKY[n + i] = 0.0;
         if (Nfft4GPSolverFgmres(K11_mat, n, matvec, precond_data, precond_solve, iKY, label_helper, n, n, atol, tol, &rel_res, &rel_res_v, &niter,
                                 print_level))
            return -1;
         for (j = 0; j < n; j++)
            dot += label_helper[j] * iKY[j];
         std_predict[i] = sqrt(fabs(K22i - dot));
         free(rel_res_v);
         rel_res_v = NULL;
      }
      *std_predictp = std_predict;
   }
   free(iKY);
   free(label_helper);
   return 0;
}
