/* Multi-GPU plumbing (one process per GPU): NCCL is bound lazily with dlopen so that single-GPU users carry
 * no NCCL dependency.  The only data-path collective is ONE allreduce(sum, fp64) per matvec:
 *   mode 1 (shard by feature window): allreduce of the partial output vector, n doubles;
 *   mode 2 (shard by point block)   : allreduce of the concatenated oversampled grids (P * 64^d doubles).
 *  */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nfft4gp_b200.h"
#include "b200_dev.h"
#include "b200_host.h"

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_comm_init)(nccl_comm *, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm, void *);
typedef const char *(*fn_errstr)(int);

static void *g_nccl = NULL;
static fn_get_uid p_get_uid;
static fn_comm_init p_comm_init;
static fn_comm_destroy p_comm_destroy;
static fn_allreduce p_allreduce;
static fn_errstr p_errstr;
static nccl_comm g_comm = NULL;
static int g_rank = 0, g_size = 1;

static int nccl_bind(void)
{
   const char *names[] = {"libnccl.so.2", "libnccl.so", NULL};
   int i;
   if (g_nccl)
      return 0;
   for (i = 0; names[i] && !g_nccl; i++)
      g_nccl = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
   if (!g_nccl)
   {
      printf("Error: nfft4gp_b200 could not load NCCL (%s).\n", dlerror());
      return -1;
   }
   p_get_uid = (fn_get_uid)dlsym(g_nccl, "ncclGetUniqueId");
   p_comm_init = (fn_comm_init)dlsym(g_nccl, "ncclCommInitRank");
   p_comm_destroy = (fn_comm_destroy)dlsym(g_nccl, "ncclCommDestroy");
   p_allreduce = (fn_allreduce)dlsym(g_nccl, "ncclAllReduce");
   p_errstr = (fn_errstr)dlsym(g_nccl, "ncclGetErrorString");
   if (!p_get_uid || !p_comm_init || !p_comm_destroy || !p_allreduce)
   {
      printf("Error: nfft4gp_b200: NCCL symbols missing.\n");
      return -1;
   }
   return 0;
}

int Nfft4GPB200CommGetUniqueId(void *id128)
{
   nccl_uid id;
   int rc;
   if (nccl_bind())
      return -1;
   rc = p_get_uid(&id);
   if (rc)
   {
      printf("Error: ncclGetUniqueId: %s\n", p_errstr ? p_errstr(rc) : "?");
      return -1;
   }
   memcpy(id128, &id, sizeof(id));
   return 0;
}

int Nfft4GPB200CommInit(int rank, int nranks, const void *id128)
{
   nccl_uid id;
   int rc;
   if (b200_init(-1) || nccl_bind())
      return -1;
   memcpy(&id, id128, sizeof(id));
   rc = p_comm_init(&g_comm, nranks, id, rank);
   if (rc)
   {
      printf("Error: ncclCommInitRank: %s\n", p_errstr ? p_errstr(rc) : "?");
      return -1;
   }
   g_rank = rank;
   g_size = nranks;
   return 0;
}

int Nfft4GPB200CommFinalize(void)
{
   if (g_comm)
   {
      b200_sync();
      p_comm_destroy(g_comm);
      g_comm = NULL;
   }
   g_rank = 0;
   g_size = 1;
   return 0;
}

int b200_comm_rank(void) { return g_rank; }
int b200_comm_size(void) { return g_size; }

int b200_comm_allreduce_sum(double *d_buf, size_t count)
{
   int rc;
   if (g_size == 1)
      return 0;
   if (!g_comm)
   {
      printf("Error: nfft4gp_b200: sharded matvec without Nfft4GPB200CommInit.\n");
      return -1;
   }
   rc = p_allreduce(d_buf, d_buf, count, 8 /* ncclFloat64 */, 0 /* ncclSum */, g_comm, b200_stream_handle());
   if (rc)
   {
      printf("Error: ncclAllReduce: %s\n", p_errstr ? p_errstr(rc) : "?");
      return -1;
   }
   return 0;
}

/* window w is computed by this rank?  (mode 1: round robin over ranks) */
int b200_comm_owns_window(const b200_additive *A, int w)
{
   if (A->shard_mode != 1 || g_size == 1)
      return 1;
   return (w % g_size) == g_rank;
}

/* local pointer arrays of the windows this rank owns (mode 1); called once the windows exist on the device */
int b200_comm_build_local_windows(b200_additive *A)
{
   int P = A->P, i, nl = 0, rc;
   const double **fp;
   const int **ip;
   if (A->shard_mode != 1 || g_size == 1 || A->d_fptr_local)
      return 0;
   fp = (const double **)b200_xmalloc(sizeof(double *) * (size_t)P);
   ip = (const int **)b200_xmalloc(sizeof(int *) * (size_t)P);
   for (i = 0; i < P; i++)
      if (b200_comm_owns_window(A, i))
      {
         fp[nl] = A->win[i]->dev.f;
         ip[nl] = A->win[i]->dev.f_orig ? NULL : A->win[i]->dev.inv;
         nl++;
      }
   A->nlocal = nl;
   A->d_fptr_local = (const double **)b200_malloc(sizeof(double *) * (size_t)(nl ? nl : 1));
   A->d_invptr_local = (const int **)b200_malloc(sizeof(int *) * (size_t)(nl ? nl : 1));
   if (!A->d_ypart)
      A->d_ypart = (double *)b200_malloc(sizeof(double) * (size_t)A->n);
   rc = b200_h2d((void *)A->d_fptr_local, fp, sizeof(double *) * (size_t)nl) ||
        b200_h2d((void *)A->d_invptr_local, ip, sizeof(int *) * (size_t)nl) || b200_sync();
   free(fp);
   free(ip);
   return rc ? -1 : 0;
}

/* Must be called BEFORE the first kernel build of the handle (the point tables are built per shard). */
int Nfft4GPB200SetSharding(void *handle, int mode)
{
   b200_additive *A = b200_additive_from_handle(handle);
   if (!A || mode < 0 || mode > 2)
   {
      printf("Error: Nfft4GPB200SetSharding: bad handle or unknown mode %d.\n", mode);
      return -1;
   }
   if (A->built)
   {
      printf("Error: Nfft4GPB200SetSharding must precede the first kernel build of the handle.\n");
      return -1;
   }
   A->shard_mode = mode;
   return 0;
}

/* y = beta*y + sum_w ff*(f_w + mu*s*x): single GPU in one kernel; window-sharded = local partial, ONE allreduce, axpy */
int b200_additive_epilogue(b200_additive *A, int n, const double *d_x, double s, double ff, double mu, double beta, double *d_y)
{
   if (A->shard_mode == 1 && g_size > 1)
   {
      if (b200k_combine(n, A->nlocal, A->d_fptr_local, A->d_invptr_local, d_x, s, ff, mu, 0.0, A->d_ypart))
         return -1;
      if (b200_comm_allreduce_sum(A->d_ypart, (size_t)n))
         return -1;
      if (beta == 0.0)
         return b200_d2d(d_y, A->d_ypart, sizeof(double) * (size_t)n);
      if (beta != 1.0 && b200k_scale(n, beta, d_y))
         return -1;
      return b200k_axpy(n, 1.0, A->d_ypart, d_y);
   }
   return b200k_combine(n, A->P, A->d_fptr, A->d_invptr, d_x, s, ff, mu, beta, d_y);
}
