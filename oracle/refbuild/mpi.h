/* TEST INFRASTRUCTURE ONLY (oracle/). Single-rank MPI stand-in used when the
 * reference sources under /root/reference are compiled, unmodified and where
 * they lie, into oracle/_ref/libref.so.  No MPI exists in this image; the
 * reference's serial solvers never communicate, and its "MPI" solvers run as
 * rank 0 of 1 with every neighbour == MPI_PROC_NULL.  Written from the MPI
 * standard's signatures; nothing here comes from the reference. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD        0
#define MPI_PROC_NULL        (-2)
#define MPI_SUCCESS           0
#define MPI_INT               1
#define MPI_DOUBLE            2
#define MPI_C_DOUBLE_COMPLEX  3

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm old, int ndims, const int dims[], const int periods[],
                    int reorder, MPI_Comm *cart);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *src, int *dst);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[]);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype,
                    MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag,
              MPI_Comm comm, MPI_Request *req);
int MPI_Irecv(void *buf, int count, MPI_Datatype type, int src, int tag,
              MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype stype, int dest, int stag,
                 void *rbuf, int rcount, MPI_Datatype rtype, int src, int rtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm,
             MPI_Status *status);
#endif
