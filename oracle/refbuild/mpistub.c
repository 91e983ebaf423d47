/* TEST INFRASTRUCTURE ONLY (oracle/). Rank 0 of 1; all point-to-point calls
 * address MPI_PROC_NULL and are therefore no-ops, as the MPI standard says. */
#include "mpi.h"

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
int MPI_Dims_create(int nnodes, int ndims, int dims[])
{ (void)nnodes; for (int d = 0; d < ndims; ++d) dims[d] = 1; return MPI_SUCCESS; }
int MPI_Cart_create(MPI_Comm old, int ndims, const int dims[], const int periods[],
                    int reorder, MPI_Comm *cart)
{ (void)old; (void)ndims; (void)dims; (void)periods; (void)reorder; *cart = 0; return MPI_SUCCESS; }
int MPI_Cart_shift(MPI_Comm c, int direction, int disp, int *src, int *dst)
{ (void)c; (void)direction; (void)disp; *src = MPI_PROC_NULL; *dst = MPI_PROC_NULL; return MPI_SUCCESS; }
int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int coords[])
{ (void)c; (void)rank; for (int d = 0; d < maxdims; ++d) coords[d] = 0; return MPI_SUCCESS; }
int MPI_Type_vector(int count, int bl, int stride, MPI_Datatype old, MPI_Datatype *newtype)
{ (void)count; (void)bl; (void)stride; (void)old; *newtype = 100; return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *t) { (void)t; return MPI_SUCCESS; }
int MPI_Isend(const void *b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)dest; (void)tag; (void)c; *r = 0; return MPI_SUCCESS; }
int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *r)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; *r = 0; return MPI_SUCCESS; }
int MPI_Wait(MPI_Request *r, MPI_Status *s) { (void)r; (void)s; return MPI_SUCCESS; }
int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int dest, int stag,
                 void *rb, int rn, MPI_Datatype rt, int src, int rtag, MPI_Comm c, MPI_Status *s)
{ (void)sb; (void)sn; (void)st; (void)dest; (void)stag; (void)rb; (void)rn; (void)rt;
  (void)src; (void)rtag; (void)c; (void)s; return MPI_SUCCESS; }
int MPI_Send(const void *b, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dest; (void)tag; (void)c; return MPI_SUCCESS; }
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *s)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)s; return MPI_SUCCESS; }
