/* ============================================================================
 * ippl_b200.h -- C ABI of the B200-native particle-mesh hot path.
 *
 * Drop-in boundary for IPPL's PIC hot path.  The reference has no FFI: its
 * boundary is the C++ template API (ippl::scatter / ippl::gather / ParticleBase::update /
 * BareField::accumulateHalo / fillHalo).  The host-side C++ facade in include/ippl/ keeps that
 * API and forwards to the entry points below; each entry point cites the reference interface it
 * replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; ipplb_last_error() gives the text.
 *    Nothing throws across this boundary.  There is NO CPU fallback: without a CUDA device every
 *    compute entry point fails with IPPLB_ERR_NO_DEVICE.
 *  - all array pointers are DEVICE pointers owned by the caller unless the name ends in _host.
 *  - one ipplb_ctx per GPU / rank.  All work is enqueued on the context's stream; calls on one
 *    context must not be concurrent.  ipplb_sync() waits for the stream (IpplTimings fences).
 *  - particles are SoA fp64: x,y,z / px,py,pz / q (the reference stores AoS Vector<double,3>,
 *    src/Particle/ParticleAttrib.h:33-277; the facade presents view(i)[d] over SoA).
 *  - fields are ghosted, x fastest: idx = i + ex*(j + ey*k), ex = nl[0] + 2*nghost, ncomp doubles
 *    per cell interleaved (rho: 1, E: 3 == Vector<double,3> per cell as in src/Field/BareField.h).
 * ========================================================================== */
#ifndef IPPL_B200_H
#define IPPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipplb_ctx ipplb_ctx;
typedef struct ipplb_layout ipplb_layout;   /* host-only: rank boxes + neighbour tables */
typedef struct ipplb_loop ipplb_loop;       /* in-process rank group (tests): see the multi-rank section */
typedef struct ipplb_poisson ipplb_poisson; /* cuFFT periodic Poisson solver (non-owned stage) */
typedef struct ipplb_bins ipplb_bins;       /* cell-ordered particle store behind the fused step (see below) */

enum {
    IPPLB_OK = 0,
    IPPLB_ERR_ARG = 1,
    IPPLB_ERR_CUDA = 2,
    IPPLB_ERR_NO_DEVICE = 3,
    IPPLB_ERR_NCCL = 4,
    IPPLB_ERR_CAPACITY = 5,
    IPPLB_ERR_CUFFT = 6
};

/* Local view of the mesh: UniformCartesian (src/Meshes/UniformCartesian.h) + the rank's box of the
 * FieldLayout (src/FieldLayout/FieldLayout.h, getLocalNDIndex) + BareField's ghost width. */
typedef struct ipplb_mesh {
    int ng[3];        /* global cells per dim */
    int first[3];     /* first global cell of the local box */
    int nl[3];        /* local cells per dim */
    int nghost;       /* 1 in the reference (src/Field/BareField.hpp:100) */
    double origin[3];
    double h[3];
} ipplb_mesh;

/* SoA particle bundle handed to the fused / sorting entry points. q may be NULL when the charge is
 * uniform (alpine: q = Q/totalP, demos/alpine/LandauDampingManager.h:246): q_scalar is used. */
typedef struct ipplb_particles {
    double* x; double* y; double* z;
    double* px; double* py; double* pz;
    double* q;
    double q_scalar;
    long n;         /* local particle count */
    long capacity;  /* allocated elements per array */
} ipplb_particles;

/* Push variants fused into the gather. */
enum { IPPLB_PUSH_LEAPFROG = 0, IPPLB_PUSH_PENNING = 1 };
typedef struct ipplb_push {
    int kind;
    double dt;
    /* which sub-steps run, in reference order (demos/alpine/LandauDampingManager.h:265-320):
     *   gather E at R -> kick2 (end of step n) -> kick1 (start of step n+1) -> drift -> periodic BC */
    int do_kick2, do_kick1, do_drift, do_bc;
    /* Penning trap only (demos/alpine/PenningTrapManager.h:56-74, 242-333) */
    double origin[3], length[3], V0, alpha, Bext, DrInv;
} ipplb_push;

const char* ipplb_last_error(void);
const char* ipplb_version(void);

/* ---- context ---------------------------------------------------------------------------- */
/* stream: the cudaStream_t to enqueue on (NULL = the CUDA default stream, which is what a host
 * framework such as torch uses unless told otherwise).  create_stream != 0: ignore `stream` and let the
 * context create and own a non-blocking stream. */
int ipplb_ctx_create(ipplb_ctx** out, int device, void* stream, int create_stream);
int ipplb_ctx_destroy(ipplb_ctx* ctx);
int ipplb_sync(ipplb_ctx* ctx);
void* ipplb_ctx_stream(ipplb_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long ipplb_launch_count(ipplb_ctx* ctx);

/* ---- scatter / gather (replaces ParticleAttrib::scatter / gather) --------------------------- */
/* ippl::scatter(attrib, f, pp[, policy, hash]) kernel part, src/Particle/ParticleAttrib.hpp:132-184 +
 * src/Interpolation/CIC.hpp:26-45.  Deposits particles [begin,end) (hash == NULL) or hash[begin..end)
 * into the ghosted field rho (+=).  Order-independent correctness; NO halo accumulate (chain
 * ipplb_halo_accumulate_periodic / ipplb_halo_exchange like BareField::accumulateHalo does). */
int ipplb_scatter_cic(ipplb_ctx* ctx, const ipplb_mesh* mesh, long begin, long end, const double* x,
                      const double* y, const double* z, const double* q, double q_scalar,
                      const int* hash, double* rho);
/* Same deposit for particles that are sorted by cell with cell_offsets[ncells+1] (from
 * ipplb_sort_by_cell): one thread per (cell, stencil node), register accumulation, one
 * reduction per node -- the fast path. */
int ipplb_scatter_cic_sorted(ipplb_ctx* ctx, const ipplb_mesh* mesh, long n, const double* x,
                             const double* y, const double* z, const double* q, double q_scalar,
                             const int* cell_offsets, double* rho);
/* ippl::gather(attrib, f, pp, addToAttribute) kernel part, ParticleAttrib.hpp:193-246 + CIC.hpp:47-66.
 * field: ghosted, ncomp (1 or 3) interleaved comps; out[c] arrays of n. NO halo fill. */
int ipplb_gather_cic(ipplb_ctx* ctx, const ipplb_mesh* mesh, long n, const double* x, const double* y,
                     const double* z, const double* field, int ncomp, double* const* out_host_ptrs,
                     int add_to_attribute);
/* Fused gather + push: E never round-trips through HBM (north_star).  efield: ghosted AoS-3 with
 * valid halo.  Reads and rewrites x,y,z,px,py,pz in place. */
int ipplb_gather_push(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push,
                      ipplb_particles* p, const double* efield);
/* How ipplb_gather_cic (3 components) and ipplb_gather_push read the field: 1 (default) one 8-byte load per node and
 * component (24 per particle); 2 the two x-neighbours of a stencil row with 16-byte loads (14 per particle) -- fewer load
 * wavefronts on unordered particles, where every lane touches its own sectors.  Bit-identical results.  Falls back to 1
 * when the field pointer is not 16-byte aligned. */
int ipplb_ctx_set_gather_variant(ipplb_ctx* ctx, int variant);

/* ---- unfused particle ops (arbitrary driver expressions) ---------------------------------- */
/* y[i] = y[i] + a * x[i]  (ParticleAttrib::operator=(Expression), ParticleAttrib.hpp:118-130, for
 * P = P - 0.5*dt*E (a = -0.5*dt; identical rounding to P - c*E) and R = R + dt*P) */
int ipplb_axpy(ipplb_ctx* ctx, long n, double a, const double* x, double* y);
/* PeriodicBC per dimension, src/Particle/ParticleBC.h:73-76 via ParticleLayout::applyBC
 * (src/Particle/ParticleLayout.hpp:34-74); lo/hi = global region bounds; mask bit d = dim d periodic */
int ipplb_apply_periodic_bc(ipplb_ctx* ctx, long n, double* x, double* y, double* z,
                            const double lo[3], const double hi[3], int mask);
/* PenningTrap Kick1 / Kick2 as separate passes (PenningTrapManager.h:256-272, 313-333) */
int ipplb_penning_kick(ipplb_ctx* ctx, int which, const ipplb_push* push, long n, const double* x,
                       const double* y, const double* z, double* px, double* py, double* pz,
                       const double* ex, const double* ey, const double* ez);

/* ---- cell sort (B200-native; replaces the role of Interpolation/Binning.h bin_sort) -------- */
/* Counting sort by cell key (integer, bit-exact: key from index = (int)((x-origin)*invdx + 0.5),
 * the same truncation scatter/gather use).  Out of place: in -> out (arrays must not alias).
 * cell_offsets (device, ncells+1 ints, cells of the local box grown by one layer on the upper side,
 * see ipplb_sort_ncells) receives the start of every cell in the sorted order. */
long ipplb_sort_ncells(const ipplb_mesh* mesh);
int ipplb_sort_by_cell(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_particles* in,
                       ipplb_particles* out, int* cell_offsets);

/* ---- field ops ------------------------------------------------------------------------------ */
int ipplb_field_fill(ipplb_ctx* ctx, double* field, long count, double value);
/* BareField::sum over interior cells (src/Field/BareField.hpp:224-240); result written to *out_host
 * after a stream sync (the reference's sum is host-synchronous too). */
int ipplb_field_sum(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, double* out_host);
/* AlpineManager::getDensity, demos/alpine/AlpineManager.h:225-245: v = v / cell_volume - shift on the
 * interior (two IEEE ops in that order, like the reference's two expression kernels). */
int ipplb_field_density(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, double cell_volume,
                        double shift);
/* HaloCells::applyPeriodicSerialDim (src/Field/HaloCells.hpp:297-336): in-rank periodic wrap for the
 * dims in serial_mask (bit d), cascading d = 0,1,2 like the reference.  accumulate: ghost += into the
 * opposite interior layer; fill: ghost = opposite interior layer. */
int ipplb_halo_accumulate_periodic(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, int ncomp,
                                   int serial_mask);
int ipplb_halo_fill_periodic(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, int ncomp,
                             int serial_mask);
/* Ex field energy + max norm of demos/alpine/LandauDampingManager.h:339-366 (interior cells, comp 0):
 * out_host[0] = sum(Ex^2), out_host[1] = max|Ex| */
int ipplb_field_ex_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* efield,
                         double* out_host);

/* ---- periodic FFT Poisson solve (NON-OWNED stage, cuFFT; timed separately) ----------------- */
/* FFTPeriodicPoissonSolver::solve with output_type GRAD, src/PoissonSolvers/
 * FFTPeriodicPoissonSolver.hpp:53-169 (forward scaled 1/N, Nyquist and DC zeroed, unscaled inverse).
 * Single-GPU: mesh must be the whole domain.  rho: ghosted scalar field (interior read; like the
 * reference, rho's interior is clobbered).  efield: ghosted AoS-3, interior written (halo NOT filled). */
int ipplb_poisson_create(ipplb_ctx* ctx, const ipplb_mesh* mesh, ipplb_poisson** out);
/* Multi-rank variant: a REPLICATED solve over NVSwitch.  ipplb_poisson_solve then gathers every rank's rho interior on
 * every rank (one grouped ncclBroadcast per rank box, so ORB boxes of unequal size work), solves the whole domain with
 * the same cuFFT plan on every GPU and keeps this rank's box of E (halo NOT filled: chain ipplb_halo_exchange).  Replaces
 * heFFTe's distributed transposes (src/FFT/Transform/RC.h) for grids that fit one GPU: 256^3 is 1.1 GB of work space,
 * 512^3 is 8.6 GB.  `layout` is the current FieldLayout; needs ipplb_comm_init.  With one rank it equals the above. */
int ipplb_poisson_create_dist(ipplb_ctx* ctx, const ipplb_layout* layout, const double origin[3], const double h[3],
                              ipplb_poisson** out);
int ipplb_poisson_solve(ipplb_poisson* s, double* rho, double* efield);
int ipplb_poisson_destroy(ipplb_poisson* s);
/* Multi-rank variant 2: the SLAB-DECOMPOSED solve -- what heFFTe does for the reference (src/FFT/FFT.hpp:118-193: fft3d_r2c
 * between the FieldLayout boxes; src/FFT/Transform/RC.h): boxes -> z-slabs (2-D real-to-complex transforms of whole planes)
 * -> y-slabs (1-D transforms along z, k-space multipliers, 1-D inverses) -> z-slabs (2-D inverses, three gradient
 * components) -> boxes.  Four message exchanges per solve (NCCL send / recv groups; device copies inside an in-process rank
 * group), every rank transforms 1/N of the domain and holds ~(1 + 3 + 3 + 4) / N of it as work space.  ipplb_poisson_solve
 * and _destroy work on the handle as usual; ipplb_loop_poisson_solve drives the solvers of all ranks of an in-process
 * group.  `layout` may be any box layout (ORB).  E's halo is NOT filled. */
int ipplb_poisson_create_slab(ipplb_ctx* ctx, const ipplb_layout* layout, const double origin[3], const double h[3],
                              ipplb_poisson** out);
int ipplb_loop_poisson_solve(ipplb_loop* loop, ipplb_poisson* const* solvers, double* const* rho, double* const* efield);
/* The host-side plan behind it (no GPU needed): the sub-box copies and messages of the four phases, for checks.
 * info[16] = nranks, rank, ng[3], nx/2+1, z-slab [zs, ze), y-slab [ys, ye), doubles in REAL, SPEC2D, SPECZ, SEND, RECV, nghost.
 * rows: 16 longs per row.  which = 0 / 2 (copies before / after the exchange): src_buf, dst_buf, src_off, dst_off, src
 * strides[3], dst strides[3], extents[3], elem (1 real, 2 complex; offsets and strides in elements); buffers 0 rho, 1 E,
 * 2 REAL, 3 SPEC2D, 4 SPECZ, 5 SEND, 6 RECV.  which = 1 (messages): peer, send offset, send count, recv offset, recv count
 * (doubles). */
typedef struct ipplb_slabplan ipplb_slabplan;
int ipplb_slabplan_create(const ipplb_layout* layout, int rank, ipplb_slabplan** out);
int ipplb_slabplan_info(const ipplb_slabplan* p, long info[16]);
int ipplb_slabplan_rows(const ipplb_slabplan* p, int phase, int which, long* rows, int max_rows, int* nrows);
int ipplb_slabplan_destroy(ipplb_slabplan* p);

/* ---- layout (host only; FieldLayout / Partitioner / RegionLayout) ---------------------------- */
/* FieldLayout(comm, domain, decomp, isAllPeriodic, nghost) for `nranks` ranks,
 * src/FieldLayout/FieldLayout.hpp:76-134 + src/Partition/Partitioner.hpp:15-123. */
int ipplb_layout_create(ipplb_layout** out, const int ng[3], const int is_parallel[3], int nranks,
                        int periodic, int nghost);
/* FieldLayout::updateLayout(domains) (ORB repartition): boxes[nranks][6] = lo[3], hi[3] inclusive */
int ipplb_layout_set_boxes(ipplb_layout* l, const int* boxes);
int ipplb_layout_destroy(ipplb_layout* l);
int ipplb_layout_nranks(const ipplb_layout* l);
/* boxes_out[nranks][6] = lo[3], hi[3] inclusive global cell indices (getLocalNDIndex(rank)) */
int ipplb_layout_boxes(const ipplb_layout* l, int* boxes_out);
/* getNeighbors / getNeighborsSendRange / getNeighborsRecvRange of rank `rank`, flattened in component
 * order: out[i][14] = comp, peer, send lo[3], send hi[3], recv lo[3], recv hi[3] (hi exclusive, local
 * ghosted indices).  Returns the entry count (writes at most max_entries). */
int ipplb_layout_neighbors(const ipplb_layout* l, int rank, int* out, int max_entries);
/* RegionLayout regions: regions_out[nranks][6] = min[3], max[3] (src/Region/RegionLayout.hpp:68-98) */
int ipplb_layout_regions(const ipplb_layout* l, const double origin[3], const double h[3],
                         double* regions_out);
/* fills *mesh for rank `rank` */
int ipplb_layout_mesh(const ipplb_layout* l, int rank, const double origin[3], const double h[3],
                      ipplb_mesh* mesh);

/* ---- multi-GPU (NCCL over NVLink; one process per GPU) ---------------------------------------- */
#define IPPLB_NCCL_ID_BYTES 128
int ipplb_nccl_unique_id(char id_out[IPPLB_NCCL_ID_BYTES]);
int ipplb_comm_init(ipplb_ctx* ctx, int rank, int nranks, const char id[IPPLB_NCCL_ID_BYTES]);
/* binds the decomposition: builds the device-side halo plan (all 26 components batched) and the
 * ownership tables used by ipplb_update. */
int ipplb_ctx_set_layout(ipplb_ctx* ctx, const ipplb_layout* l, const double origin[3],
                         const double h[3]);
/* BareField::accumulateHalo / fillHalo (src/Field/BareField.hpp:152-172): inter-rank exchange of all
 * neighbour components (HaloCells::exchangeBoundaries, src/Field/HaloCells.hpp:109-242) as ONE batched
 * pack kernel, ONE grouped ncclSend/ncclRecv, ONE batched unpack kernel, then the in-rank periodic wrap
 * for un-split dims.  mode: 0 fill, 1 accumulate. */
int ipplb_halo_exchange(ipplb_ctx* ctx, double* field, int ncomp, int mode);
/* ParticleSpatialLayout::update (src/Particle/ParticleSpatialLayout.hpp:115-314): periodic BC,
 * ownership (bit-exact region test, :316-330, 372-395), count exchange, SoA migration, compaction.
 * p->n is updated; arrays must have capacity for the arrivals (IPPLB_ERR_CAPACITY otherwise).
 * sent_host / recv_host (may be NULL): per-rank counts [nranks] for parity checks. */
int ipplb_update(ipplb_ctx* ctx, ipplb_particles* p, long* sent_host, long* recv_host);
/* The same in two collective halves, for callers that own growable arrays (the facade's ParticleAttrib, which the
 * reference grows on receive, src/Particle/ParticleBase.hpp:300-393): ipplb_update_plan locates the particles and
 * exchanges the counts (no particle moves); *n_after_host = this rank's count after the migration, so the caller can
 * reserve; ipplb_update_commit then packs, exchanges and unpacks.  Errors are COLLECTIVE: every rank's count, capacity
 * and outcome travel with the count exchange, so when one rank cannot hold its arrivals every rank returns
 * IPPLB_ERR_CAPACITY (from plan: reserve and call commit; from commit: fatal) and none is left waiting in a receive. */
int ipplb_update_plan(ipplb_ctx* ctx, ipplb_particles* p, long* n_after_host, long* sent_host, long* recv_host);
int ipplb_update_commit(ipplb_ctx* ctx, ipplb_particles* p);
/* sum over ranks of one double / one long (rho.sum(), particle count: AlpineManager.h:169, 212) */
int ipplb_allreduce_sum_f64(ipplb_ctx* ctx, double* value_host);
int ipplb_allreduce_sum_i64(ipplb_ctx* ctx, long* value_host);
/* max over ranks of one double (the dumps' max norms: Comm->reduce(..., std::greater<double>()), LandauDampingManager.h:360-366) */
int ipplb_allreduce_max_f64(ipplb_ctx* ctx, double* value_host);

/* ---- all ranks of a small job in ONE process on ONE device (no NCCL) ------------------------------------------
 * The multi-rank path is written as local phases (kernels of one rank) around a transport step.  ipplb_loop_* drive
 * every rank of an in-process group through the SAME phases -- same kernels, same tables, same host logic as the NCCL
 * entry points above -- with device-to-device copies (and, for the bucketed store, plain pointers into the other
 * contexts' inboxes) as the transport.  This is how ownership (ParticleSpatialLayout.hpp:316-330, 372-395), the
 * send / receive / compaction of update (:150-314, ParticleBase.hpp:175-393) and HaloCells::exchangeBoundaries
 * (HaloCells.hpp:109-242) are held to the oracle on a single GPU.  ctxs[r] becomes rank r of nranks; bind each with
 * ipplb_ctx_set_layout afterwards.  Arrays below are indexed by rank. */
int ipplb_loop_create(ipplb_loop** out, ipplb_ctx* const* ctxs, int nranks);
int ipplb_loop_destroy(ipplb_loop* loop);
int ipplb_loop_halo_exchange(ipplb_loop* loop, double* const* fields, int ncomp, int mode);
/* parts[nranks]; sent_host / recv_host (may be NULL): [nranks][nranks] counts, row r = rank r's */
int ipplb_loop_update(ipplb_loop* loop, ipplb_particles* parts, long* sent_host, long* recv_host);
int ipplb_loop_migrate_connect(ipplb_loop* loop, long seg_cap);
/* bins[nranks], cur[nranks], rho[nranks] (rho or its entries may be NULL) */
int ipplb_loop_bins_migrate(ipplb_loop* loop, ipplb_bins* const* bins, ipplb_particles* cur, double* const* rho);

/* ---- cell-ordered particle store + fused single-pass PIC step (the B200-first path) ------------------ */
/* ipplb_bins keeps the particles of one rank grouped in per-tile buckets (tile = 4x4x4 key cells, key =
 * index - first of the CIC index (int)((x-origin)*invdx+0.5), the same truncation scatter/gather use).
 * Bucket t owns slots [start[t], start[t]+cap[t]) of the six SoA arrays and holds count[t] particles in
 * runs that are sorted by cell; caps carry a few per cent of slack so ONE pass over the particles can push
 * them, re-bin them and write them straight into next step's buckets (no counting pass, no separate sort).
 * Particles that do not fit (or arrive from other ranks) live in an unsorted tail after the last bucket.
 * The tables live in device memory and are re-planned on the device after every step: no host sync.
 * This is storage behind ParticleAttrib (src/Particle/ParticleAttrib.h:33-277): ipplb_bins_compact gives
 * the contiguous [0,n) view the reference API exposes. */
enum {
    IPPLB_FLAG_EXIT_OVERFLOW = 1,  /* more leavers than exit_cap (leavers beyond it were dropped) */
    IPPLB_FLAG_CAPACITY      = 2,  /* arrays too small for buckets + tail (particles were dropped) */
    IPPLB_FLAG_INTERNAL      = 4,  /* invariant violated (bug) */
    IPPLB_FLAG_SLACK_SCALED  = 8   /* informational: bucket slack was reduced to fit the capacity */
};
/* capacity: elements per SoA array of BOTH particle bundles handed to build/step (>= ~1.25 n). */
int ipplb_bins_create(ipplb_ctx* ctx, const ipplb_mesh* mesh, long capacity, ipplb_bins** out);
int ipplb_bins_destroy(ipplb_bins* bins);
/* Counting sort of `in` (contiguous, any order, in->n particles; uniform charge) into buckets of `out`
 * (cell-sorted inside every bucket).  in and out must not alias. */
int ipplb_bins_build(ipplb_ctx* ctx, ipplb_bins* bins, const ipplb_particles* in, ipplb_particles* out);
/* How ipplb_bins_build places the particles: 1 (default) per-cell positions inside the bucket, one atomic per particle;
 * 2 arrival order inside the bucket (the order the fused step itself maintains for arrivals), one atomic per run of
 * same-tile lanes, sequential write position per tile so that the 8-byte stores merge in L2.  Same tables either way.
 * The reference's counterpart is its counting-sort binning, src/Interpolation/Binning.h:110-114. */
int ipplb_bins_set_build_variant(ipplb_bins* bins, int variant);
/* One fused step over the bucketed particles `cur`, written re-bucketed into `nxt` (the caller swaps the
 * two bundles afterwards): per particle gather E (tile of E staged in shared memory) -> kick, kick, drift,
 * periodic BC (ipplb_push, bit-identical to ipplb_gather_push) -> shared-memory binning by new cell ->
 * coalesced store into next step's buckets -> charge deposit from the sorted shared-memory copy
 * (rho +=; caller zeroes rho and chains the halo accumulate).  Asynchronous on the context's stream.
 * Input is streamed with bulk async copies (TMA, cp.async.bulk + mbarrier) by a producer warp.
 * Particles that left the rank's region (multi-GPU: reference ownership test, ParticleSpatialLayout.hpp:
 * 316-330) are not deposited: each is appended as one record (x,y,z,px,py,pz) to its destination rank's segment of
 * exit_buf[nranks][exit_cap / nranks][6] (16-byte aligned; one rank: exit_buf[exit_cap][6]).
 * One rank that owns the WHOLE periodic domain (nl == ng in every dimension, no region test, periodic BC in the push): the
 * step aliases ghost nodes to the opposite interior layer itself -- what HaloCells::applyPeriodicSerialDim does in two
 * extra passes (src/Field/HaloCells.hpp:297-336).  efield's ghost layers are then not read (no fillHalo needed) and rho's
 * ghost layers receive nothing (a chained accumulateHalo adds zeros).
 * Replaces, for one step: ParticleAttrib::operator= x3 (ParticleAttrib.hpp:118-130), applyBC
 * (ParticleLayout.hpp:34-74), gather (:193-246) and scatter (:132-184). */
int ipplb_bins_step(ipplb_ctx* ctx, ipplb_bins* bins, const ipplb_push* push, const ipplb_particles* cur,
                    ipplb_particles* nxt, const double* efield, double* rho, double* exit_buf,
                    int exit_cap, const double region_min[3], const double region_max[3]);
/* Synchronises the stream and reports the state after the last build/step/append: particles held
 * (buckets + tail), of which in the tail, leavers written to exit_buf by the last step, IPPLB_FLAG_* bits. */
int ipplb_bins_status(ipplb_ctx* ctx, ipplb_bins* bins, long* n_local, long* n_tail, long* n_exit,
                      int* flags);
/* Appends `count` particles (device SoA pointers src[6]) to the tail of `cur` (migration arrivals). */
int ipplb_bins_append(ipplb_ctx* ctx, ipplb_bins* bins, ipplb_particles* cur, const double* const src[6],
                      long count);
/* ParticleSpatialLayout::update for the bucketed store (src/Particle/ParticleSpatialLayout.hpp:115-314): the
 * leavers the last ipplb_bins_step wrote to exit_buf get their destination rank (reference search order incl.
 * the inclusive fallback, :372-395), are exchanged over NCCL and appended to the tail of `cur`; arrivals are
 * deposited into rho (may be NULL).  Updates cur->n; per-rank counts in sent_host / recv_host (may be NULL).
 * Synchronises the stream.  With one rank it only refreshes cur->n. */
int ipplb_bins_migrate(ipplb_ctx* ctx, ipplb_bins* bins, ipplb_particles* cur, const double* exit_buf,
                       int exit_cap, double* rho, long* sent_host, long* recv_host);
/* Peer-memory migration (the default multi-GPU path).  ipplb_migrate_connect (collective, once per communicator)
 * allocates this rank's inbox -- [nranks][seg_cap] 48-byte records, two copies for alternating steps -- and maps every
 * peer's inbox into this process (cudaIpc over NVLink / NVSwitch).  From then on ipplb_bins_step with exit_buf == NULL
 * writes each leaver STRAIGHT into segment `rank` of its destination's inbox, and ipplb_bins_migrate_async finishes the
 * update with one small all-gather of the counts (which is also the barrier behind which the records have landed) and
 * one kernel that drops the arrivals into their buckets (overflow: tail) and deposits their charge into rho (may be
 * NULL).  Nothing synchronises with the host: cur->n is NOT refreshed (ipplb_bins_status reports the device truth) and
 * errors surface as the sticky IPPLB_FLAG_* bits.  ipplb_migrate_counts (synchronises) returns the per-rank counts of
 * the last migration for checks.  seg_cap must be ONE number for the job (senders and receivers address the segments
 * with it): the ranks agree on the largest value any of them passed. */
int ipplb_migrate_connect(ipplb_ctx* ctx, long seg_cap);
int ipplb_bins_migrate_async(ipplb_ctx* ctx, ipplb_bins* bins, ipplb_particles* cur, double* rho);
int ipplb_migrate_counts(ipplb_ctx* ctx, long* sent_host, long* recv_host);
/* Contiguous copy (bucket order, then tail) of the bucketed `cur` into out[0..n). Sets out->n (syncs). */
int ipplb_bins_compact(ipplb_ctx* ctx, ipplb_bins* bins, const ipplb_particles* cur,
                       ipplb_particles* out);
/* the same reduction as ipplb_particles_kinetic over the bucketed store (buckets + tail), no compaction needed */
int ipplb_bins_kinetic(ipplb_ctx* ctx, ipplb_bins* bins, const ipplb_particles* cur, double* out_host);
/* Measurement (bench.py's roofline object): with timing on, every ipplb_bins_step brackets its fused kernel -- the kernel
 * alone, not the table planning behind it -- with CUDA events on the context's stream (up to 256 launches after the
 * last set_timing call).  ipplb_bins_kernel_ms synchronises the stream and returns the per-launch durations. */
int ipplb_bins_set_timing(ipplb_bins* bins, int on);
int ipplb_bins_kernel_ms(ipplb_ctx* ctx, ipplb_bins* bins, double* ms_out_host, int max_out, int* n_out);
/* read-only access for tests: copies start/cap/count of the current buffer to host arrays [ntiles] */
int ipplb_bins_ntiles(const ipplb_bins* bins);
int ipplb_bins_tables(ipplb_ctx* ctx, ipplb_bins* bins, int* start_host, int* cap_host, int* count_host);

/* ---- diagnostics of the alpine dumps ------------------------------------------ */
/* Field part of PenningTrapManager::dumpData (demos/alpine/PenningTrapManager.h:346-389), LandauDampingManager::
 * dumpLandau (LandauDampingManager.h:339-366) and BumponTailInstabilityManager::dumpBumponTailInstability
 * (BumponTailInstabilityManager.h:448-480) in one pass over the interior of the ghosted AoS-3 field:
 *   out_host[0..2] = sum(E_d^2), out_host[3..5] = max|E_d|, out_host[6] = sum(dot(E,E)) (= rho.sum() after
 *   rho = dot(E,E), PenningTrapManager.h:350-352).  Local to the rank (chain ipplb_allreduce_sum_f64). */
int ipplb_field_energy_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* efield, double out_host[7]);
/* norm(rho) ingredients (src/Field/BareField.hpp innerProduct/norm, p = 2): out_host[0] = sum over the interior of
 * f^2, out_host[1] = max|f| */
int ipplb_field_norm_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, double out_host[2]);
/* "Particle Kinetic Energy" reduction, PenningTrapManager.h:354-362: out_host[0] = sum_i dot(P_i, P_i) (the caller
 * multiplies by 0.5 like the reference) over contiguous arrays. */
int ipplb_particles_kinetic(ipplb_ctx* ctx, long n, const double* px, const double* py, const double* pz,
                            double* out_host);

/* ---- particle initialisation on the device ------------------------------------ */
/* Distribution<T, Dim, 2*Dim, Functions> of the alpine managers, one kind per dimension, parameters par[2d],
 * par[2d+1] (src/Random/Distribution.h:60-110):
 *   UNIFORM  cdf x, pdf 1, estimate u                                   (src/Random/UniformDistribution.h:17-26)
 *   COSINE   cdf x + (a/k) sin(k x), pdf 1 + a cos(k x), estimate u; a = par[2d], k = par[2d+1]
 *            (demos/alpine/LandauDampingManager.h:21-44, BumponTailInstabilityManager.h:23-52)
 *   NORMAL   cdf 0.5 (1 + erf((x-mu)/(sd sqrt 2))), pdf gaussian, estimate mu; mu = par[2d], sd = par[2d+1]
 *            (src/Random/NormalDistribution.h:11-28) */
enum { IPPLB_DIST_UNIFORM = 0, IPPLB_DIST_COSINE = 1, IPPLB_DIST_NORMAL = 2 };
typedef struct ipplb_dist {
    int kind[3];
    double par[6];
} ipplb_dist;
/* InverseTransformSampling(dist, rmax, rmin, rlayout, ntotal) for every rank at once (host only),
 * src/Random/InverseTransformSampling.h:48-61, 106-131: nlocal_out[r] = (size_t)(prod_d (cdf(locmax)-cdf(locmin)) /
 * prod_d (cdf(rmax)-cdf(rmin)) * ntotal), the first (ntotal - sum) ranks get one more; ubounds_out[r][6] = umin[3],
 * umax[3] = cdf of the rank's region bounds.  regions[nranks][6] = min[3], max[3] (ipplb_layout_regions). */
int ipplb_sample_counts(const ipplb_dist* dist, const double rmin[3], const double rmax[3], const double* regions,
                        int nranks, long ntotal, long* nlocal_out, double* ubounds_out);
/* InverseTransformSampling::generate, :172-244: per dimension u = drand(umin, umax), x = estimate(u), then
 * NewtonRaphson::solve (src/Random/Utility.h:27-60: while iter < 20 && |cdf(x) - u| > 1e-12: x -= (cdf(x) - u) /
 * pdf(x)).  The uniform stream is counter based (Philox4x32-10, key = seed, counter = (first_id + i, dimension)):
 * reproducible on any decomposition of the id range, unlike Kokkos::Random_XorShift64_Pool whose stream assignment
 * is backend dependent.  Writes x,y,z[0..n). */
int ipplb_sample_positions(ipplb_ctx* ctx, const ipplb_dist* dist, const double umin[3], const double umax[3],
                           uint64_t seed, long first_id, long n, double* x, double* y, double* z);
/* ippl::random::randn<T, Dim> (src/Random/Randn.h:82-94): p_d = mu[d] + sd[d] * N(0,1); normals by Box-Muller on
 * the same counter-based stream (dimensions 3, 4 of the counter).  Writes px,py,pz[0..n). */
int ipplb_sample_normal(ipplb_ctx* ctx, const double mu[3], const double sd[3], uint64_t seed, long first_id,
                        long n, double* px, double* py, double* pz);
/* rho = distR.getFullPdf(xvec) on the interior, xvec = (global cell index + 0.5) * h + origin: the weight field of the
 * first repartition (LandauDampingManager.h:188-199; Distribution.h:104-112 product of the per-dimension pdfs) */
int ipplb_field_fill_pdf(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_dist* dist, double* field);

/* ---- orthogonal recursive bisection -------------------------------------------- */
/* OrthogonalRecursiveBisection::binaryRepartition, src/Decomposition/OrthogonalRecursiveBisection.hpp:14-105, as a
 * host state machine (findCutAxis :107-116, findMedian :185-216, cutDomain :218-232) fed with the globally reduced
 * plane weights of perpendicularReduction (:118-183), which is the only device work:
 *   ipplb_orb_begin -> { ipplb_orb_next (domain + axis to reduce) -> plane sums -> ipplb_orb_cut } ... -> ipplb_orb_finish */
typedef struct ipplb_orb ipplb_orb;
int ipplb_orb_begin(ipplb_orb** out, const int ng[3], int nranks);
/* *pending = 1 while a cut is pending: dom_lo/dom_hi (inclusive global indices) and the cut axis of the next cut */
int ipplb_orb_next(ipplb_orb* orb, int dom_lo[3], int dom_hi[3], int* axis, int* pending);
/* feeds the reduced plane weights (n = length of the domain along the axis) and performs the cut */
int ipplb_orb_cut(ipplb_orb* orb, const double* reduced_host, int n);
/* boxes_out[nranks][6] = lo[3], hi[3]; *ok = 0 when a box has an axis of length 1 (the reference then keeps the old
 * layout, :93-99).  Destroys the state. */
int ipplb_orb_finish(ipplb_orb* orb, int* boxes_out, int* ok);
/* drops a state machine that will not be driven to ipplb_orb_finish (which frees it) */
int ipplb_orb_destroy(ipplb_orb* o);
/* perpendicularReduction + allreduce: out_host[k] = sum over ranks of the sum of `field`'s interior cells in plane
 * dom_lo[axis] + k of the domain (cells outside the rank's box contribute nothing). */
int ipplb_orb_plane_sums(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, int axis, const int dom_lo[3],
                         const int dom_hi[3], double* out_host);
/* the whole of binaryRepartition on the weight field (collective): new boxes into `boxes_out`; the caller applies
 * them with ipplb_layout_set_boxes + ipplb_ctx_set_layout, re-lays its fields and calls ipplb_update, like
 * LoadBalancer::updateLayout (demos/alpine/LoadBalancer.hpp:54-88). */
int ipplb_orb_repartition(ipplb_ctx* ctx, const ipplb_mesh* mesh, int nranks, const double* weight_field,
                          int* boxes_out, int* ok);

/* ---- whole-step conveniences used by bench.py / the facade ---------------------------------- */
/* One PIC step of the metric (scatter + push + gather) on resident particles, single rank:
 *   do_sort 0: gather_push -> rho = 0 -> atomic scatter -> periodic accumulate
 *   do_sort 1: gather_push -> counting sort -> rho = 0 -> sorted scatter -> periodic accumulate
 *   do_sort 2: rho = 0 -> ipplb_bins_step with periodic aliasing (p must be bucketed by `bins`; the mesh must be the
 *              whole periodic domain; efield's halo is not read, rho's ghost layers stay zero); p and scratch swap.
 * The field solve is NOT included (non-owned). */
int ipplb_pic_step(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, ipplb_particles* p,
                   ipplb_particles* scratch, int* cell_offsets, ipplb_bins* bins, const double* efield,
                   double* rho, int do_sort);
/* Same step through HOST buffers (bench.py's e2e): copies x..pz host->device, bins them, runs the fused
 * step, compacts and copies x..pz and rho back.  Host pointers should be pinned. */
int ipplb_pic_step_host(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, long n,
                        double* const host_arrays[6], double q_scalar, const double* efield_dev,
                        double* rho_host, ipplb_particles* dev, ipplb_particles* scratch,
                        ipplb_bins* bins, double* rho_dev);

/* Steady-state end-to-end step: the particles stay bucketed on the device (the reference's ParticleAttrib views are
 * device allocations too); per call the ghosted E field comes from HOST memory and the ghosted rho goes back to HOST memory
 * -- what a host-side (or non-owned) field solve exchanges with the particle path every step.  Single rank that owns the
 * whole periodic domain (ipplb_pic_step with do_sort = 2 in between).  Host pointers should be pinned.  Synchronises. */
int ipplb_pic_step_host_fields(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, ipplb_particles* p,
                               ipplb_particles* scratch, ipplb_bins* bins, const double* efield_host, double* rho_host,
                               double* efield_dev, double* rho_dev);

/* The same through a sequence of `nbatch` independent host batches (host_arrays[nbatch][6], rho_host[nbatch] or
 * NULL): upload of batch k+1, compute of batch k and download of batch k-1 overlap (three streams, two device
 * slots: dev[2], scratch[2], bins[2], rho_dev[2]).  Returns when every batch is back in host memory. */
int ipplb_pic_step_host_batches(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, long n, int nbatch,
                                double* const* host_arrays, double q_scalar, const double* efield_dev,
                                double* const* rho_host, ipplb_particles* dev, ipplb_particles* scratch,
                                ipplb_bins* const* bins, double* const* rho_dev);

#ifdef __cplusplus
}
#endif
#endif /* IPPL_B200_H */
