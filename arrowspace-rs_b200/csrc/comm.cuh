// comm.cuh -- NCCL plumbing of the row-sharded path: one process per GPU, items sharded by row, centroids /
// feature Laplacian / queries replicated.  NCCL (over NVLink 5 / NVSwitch) carries exactly what has to travel:
//   * Two-NN: the 500 sample rows (all-reduce of a gather buffer) and the per-shard two nearest distances (all-gather);
//   * clustering: the K x F centroid state down the ranks (send / recv, <= 6.1 MB) and its final broadcast;
//   * the feature Laplacian CSR (broadcast from rank 0), the lambda statistics (all-reduce of 3 doubles,
//     src/eigenmaps.rs:372-382), the per-shard top-k lists (all-gather, Q x k x 16 B per rank).
// No collective ever touches the N x F items.  libnccl.so.2 is opened at run time (dlopen): a process that already
// loaded NCCL (torch does) shares that copy, a plain C / Rust host gets the system library, and a single-GPU process
// never needs it.
#pragma once

#include "common.cuh"

struct asb_comm {
    void *nccl = nullptr;   // ncclComm_t
    int rank = 0, nranks = 1;
    bool owned = false;     // created by asb_comm_init_rank (destroyed with the handle) or borrowed from the host
};

enum AsbRedOp { ASB_RED_SUM = 0, ASB_RED_MAX = 2, ASB_RED_MIN = 3 };   // ncclRedOp_t values

bool asb_nccl_available(std::string *why);
int asb_comm_bcast_bytes(asb_ctx *ctx, asb_comm *comm, void *buf_d, size_t bytes, int root);
int asb_comm_allreduce_f64(asb_ctx *ctx, asb_comm *comm, double *buf_d, size_t count, AsbRedOp op);
int asb_comm_allreduce_i64(asb_ctx *ctx, asb_comm *comm, long long *buf_d, size_t count, AsbRedOp op);
int asb_comm_allgather_bytes(asb_ctx *ctx, asb_comm *comm, const void *send_d, void *recv_d, size_t bytes_per_rank);
int asb_comm_send_bytes(asb_ctx *ctx, asb_comm *comm, const void *buf_d, size_t bytes, int peer);
int asb_comm_recv_bytes(asb_ctx *ctx, asb_comm *comm, void *buf_d, size_t bytes, int peer);
int asb_comm_group_start(asb_ctx *ctx);
int asb_comm_group_end(asb_ctx *ctx);

// the order-dependent walk over a row-sharded dataset (cluster_replay.cu): every rank returns the FINAL state
int asb_dev_cluster_sharded(asb_ctx *ctx, asb_comm *comm, const double *rows_d, int64_t n_local, int64_t f,
                            int64_t max_clusters, double radius, double *centroids_d, int64_t *assign_d,
                            unsigned long long *sizes_d, int64_t *x_out_host);
