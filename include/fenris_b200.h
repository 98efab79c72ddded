/*
 * fenris_b200.h - C ABI of the B200-native global operator/stiffness assembly path.
 *
 * This is the drop-in boundary for ONE hot path of InteractiveComputerGraphics/fenris (reference
 * @ 7181b15): `CsrAssembler` / `CsrParAssembler` {assemble_pattern, assemble, assemble_into_csr}
 * (src/assembly/global.rs:65,124,133,206,301,314) fed by an `ElementEllipticAssembler`
 * (src/assembly/local/elliptic.rs:153-158) = { space: Mesh, operator, UniformQuadratureTable, u }.
 * The reference has no FFI of its own (it is 100 % Rust); each entry point below names the
 * reference item a Rust `-sys` shim would bind it to (see INTEGRATION.md for that shim).
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; every function returns an fb200_status
 *     (0 = OK) and never throws; details via fb200_last_error().
 *   - indices at the boundary are uint64_t (Rust `usize`), scalars are double (every reference
 *     call site instantiates T = f64); the device narrows indices to 32 bit and checks ranges.
 *   - host pointers are borrowed for the duration of the call only; the context owns all
 *     device memory.  A context is bound to one CUDA device and one stream and must not be
 *     used from two host threads at once (CsrAssembler is likewise !Sync, global.rs:30).
 *   - there is NO CPU fallback: unsupported element/operator combinations return
 *     FB200_ERR_UNSUPPORTED, a missing GPU returns FB200_ERR_CUDA.
 */
#ifndef FENRIS_B200_H
#define FENRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_VERSION 1

typedef int32_t fb200_status;
enum {
    FB200_OK = 0,
    FB200_ERR_SINGULAR_JACOBIAN = 1,   /* eyre!("Singular element Jacobian encountered"), elliptic.rs:401-404 */
    FB200_ERR_SHAPE = 2,               /* assert!s at elliptic.rs:378-391, global.rs:515-517 */
    FB200_ERR_INDEX_OOB = 3,           /* connectivity index >= num_nodes (space_impl.rs:35,44 panics) */
    FB200_ERR_COLUMN_NOT_IN_PATTERN = 4, /* "Could not find column index ...", global.rs:533 */
    FB200_ERR_UNSUPPORTED = 5,         /* no specialisation and no CPU fallback */
    FB200_ERR_CUDA = 6,
    FB200_ERR_NCCL = 7,
    FB200_ERR_STATE = 8,               /* call order violated (e.g. assemble before a pattern exists) */
    FB200_ERR_COLORING = 9,            /* adopted colours are not disjoint (DisjointSubsets::try_from..., fenris-paradis/src/lib.rs:184-220) */
    FB200_ERR_NOT_CONVERGED = 10,      /* SolveErrorKind::MaxIterationsReached (fenris-sparse/src/cg.rs:278-286) */
    FB200_ERR_INDEFINITE = 11          /* SolveErrorKind::IndefiniteOperator / IndefinitePreconditioner */
};

/* Element types = the reference's connectivity newtypes (src/connectivity.rs:182,523,607,667,912). */
enum {
    FB200_QUAD4 = 1,  /* Quad4d2Connectivity  */
    FB200_TET4 = 2,   /* Tet4Connectivity     */
    FB200_HEX8 = 3,   /* Hex8Connectivity     */
    FB200_HEX27 = 4,  /* Hex27Connectivity (geometry from the first 8 vertices, hexahedron.rs:318-335) */
    FB200_TET10 = 5,  /* Tet10Connectivity (geometry from the first 4 vertices, tetrahedron.rs:226-246) */
    FB200_HEX20 = 6   /* Hex20Connectivity, the serendipity hexahedron (hexahedron.rs:369-563; geometry from the first 8 vertices) */
};

/* Operators = EllipticContraction implementors on the path. */
enum {
    FB200_LAPLACE = 1,        /* LaplaceOperator, src/assembly/operators/laplace.rs:14,60-72; Parameters = () */
    FB200_LINEAR_ELASTIC = 2, /* MaterialEllipticOperator<LinearElasticMaterial>, fenris-solid/src/lib.rs:412-508,
                                 materials.rs:108-122; Parameters = LameParameters{mu, lambda} */
    FB200_STVK = 3,           /* MaterialEllipticOperator<StVKMaterial>, fenris-solid/src/materials.rs:355-469: state-dependent (F = I +
                                 (grad u)^T); Parameters = LameParameters; uniform quadrature tables only */
    FB200_NEO_HOOKEAN = 4     /* MaterialEllipticOperator<NeoHookeanMaterial>, materials.rs:232-353 (log J through logdet.rs:17-86); as STVK;
                                 det F <= 0 gives NaN stress / contraction and +inf energy, as in the reference */
};

/* How element contributions reach the CSR values. All three give the same sums up to fp reassociation. */
enum {
    FB200_SCATTER_ATOMIC = 0,  /* element-parallel, f64 atomic add (red.global.add.f64) */
    FB200_SCATTER_COLORED = 1, /* CsrParAssembler semantics: one launch per colour, plain read-modify-write (global.rs:322-373) */
    FB200_SCATTER_GATHER = 2   /* row-owner: each CSR value is produced by exactly one thread group, no atomics, deterministic */
};

typedef struct fb200_ctx fb200_ctx;

/* UniformQuadratureTable<T, D, Data> (src/assembly/local/quadrature_table.rs:213-298):
 * one rule + one data entry per point, shared by every element. */
typedef struct fb200_quadrature {
    int32_t num_points;
    int32_t dim;            /* reference dimension (2 or 3) */
    const double* weights;  /* [num_points] */
    const double* points;   /* [num_points * dim], point-major */
    const double* data;     /* operator Parameters per point: NULL for LAPLACE; [num_points * 2] = (mu, lambda) for LINEAR_ELASTIC
                               (UniformQuadratureTable::with_uniform_data repeats one value, quadrature_table.rs:264-266) */
} fb200_quadrature;

typedef struct fb200_operator {
    int32_t kind; /* FB200_LAPLACE | FB200_LINEAR_ELASTIC | FB200_STVK | FB200_NEO_HOOKEAN */
} fb200_operator;

/* ---- lifetime / errors ------------------------------------------------------------------ */
fb200_status fb200_create(int32_t cuda_device, fb200_ctx** out);
void fb200_destroy(fb200_ctx* ctx);
const char* fb200_status_string(fb200_status s);
/* Message of the last failure on ctx (NUL-terminated, truncated to len) and, for SINGULAR_JACOBIAN /
 * INDEX_OOB / COLUMN_NOT_IN_PATTERN, the smallest offending element index (-1 otherwise). */
fb200_status fb200_last_error(fb200_ctx* ctx, char* buf, size_t len, int64_t* element_index);
int32_t fb200_abi_version(void);

/* Adopt an external CUDA stream (e.g. torch's current stream) for all subsequent work; NULL = ctx-owned stream. */
fb200_status fb200_set_stream(fb200_ctx* ctx, void* cuda_stream);
fb200_status fb200_synchronize(fb200_ctx* ctx); /* waits, then reports deferred kernel errors (singular Jacobian ...) */
/* cudaEvent pair on the ctx stream, for callers without a CUDA binding of their own. */
fb200_status fb200_timer_begin(fb200_ctx* ctx);
fb200_status fb200_timer_end(fb200_ctx* ctx, float* milliseconds); /* synchronizes */
/* Number of kernels this library launched on ctx since creation (bench.py's gpu_launches). */
uint64_t fb200_launch_count(fb200_ctx* ctx);
/* Kernel-selection knobs (no reference counterpart; results never depend on them beyond fp reassociation).
 * "hex8_tile": elements per shared-memory tile of the Hex8 atomic scatter - 64 (default) or 0 = per-element kernel.
 * "hex8_colored_tiles": 1 (default) = FB200_SCATTER_COLORED of a Hex8 space with uniform operator data colours TILES of 64 elements instead
 *   of elements (one launch of the tile kernel per tile colour, sums inside a tile in a fixed order, colours in a fixed order: bitwise
 *   reproducible, ~5x faster than per-element colours); 0 = the caller's / fb200_color_nodes' element colours, one launch per colour.
 * "hex8_owner_stores": 1 (default) = an overwriting Hex8 tile assembly needs no zero-fill of the values: every CSR row is stored by exactly
 *   one tile (the lowest-numbered one that touches the node, which also writes the row's zeros) and the other tiles reduce into it after
 *   that tile has published its stores; 0 = zero-fill all values, store tile-complete rows only, reduce into the rest. */
fb200_status fb200_set_tuning(fb200_ctx* ctx, const char* name, int32_t value);

/* ---- the space: Mesh<f64, D, C>  (src/mesh.rs:23-40) --------------------------------------- */
/* vertices: num_nodes x d AoS (Vec<OPoint<f64,D>>); connectivity: num_elements x n row-major (Vec<C([usize; n])>).
 * Implements ElementConnectivityAssembler for the mesh (local.rs:49-75) and the FiniteElementSpace gather. */
fb200_status fb200_space_upload(fb200_ctx* ctx, int32_t element_type, uint64_t num_nodes, const double* vertices,
                                uint64_t num_elements, const uint64_t* connectivity);
/* Replace the vertex coordinates only (same sizes). */
fb200_status fb200_space_update_vertices(fb200_ctx* ctx, const double* vertices);
/* A bare ElementConnectivityAssembler (local.rs:18-47) with ragged elements (NestedVec layout:
 * element_offsets[num_elements+1], element_nodes[]), e.g. the reference's MockElementAssembler
 * (tests/unit_tests/assembly/global.rs:238-264).  Pattern and colouring only. */
fb200_status fb200_connectivity_upload(fb200_ctx* ctx, uint64_t num_nodes, uint64_t num_elements,
                                       const uint64_t* element_offsets, const uint64_t* element_nodes);
/* Multi-GPU: only elements [0, num_owned) are assembled; the rest are ghosts that complete the
 * sparsity pattern of interface rows.  Default: all owned. */
fb200_status fb200_set_num_owned_elements(fb200_ctx* ctx, uint64_t num_owned);

/* ---- CsrAssembler::assemble_pattern (global.rs:65-120 == 206-297) ------------------------- */
fb200_status fb200_assemble_pattern(fb200_ctx* ctx, int32_t solution_dim, uint64_t* num_rows, uint64_t* nnz);
/* Exactly the reference's SparsityPattern: offsets[num_rows+1], sorted col indices[nnz]. Either pointer may be NULL. */
fb200_status fb200_pattern_download(fb200_ctx* ctx, uint64_t* row_offsets, uint64_t* col_indices);
/* Use the caller's existing CsrMatrix pattern (assemble_into_csr on a matrix built elsewhere).  The pattern must be
 * node-block structured (what assemble_pattern produces); otherwise FB200_ERR_UNSUPPORTED. Missing couplings are
 * reported as FB200_ERR_COLUMN_NOT_IN_PATTERN (global.rs:533). */
fb200_status fb200_pattern_adopt(fb200_ctx* ctx, int32_t solution_dim, uint64_t num_rows, const uint64_t* row_offsets,
                                 const uint64_t* col_indices);

/* ---- color_nodes (global.rs:540-551 -> fenris-paradis/src/coloring.rs:6-70) ---------------- */
fb200_status fb200_color_nodes(fb200_ctx* ctx, uint64_t* num_colors);
/* Vec<DisjointSubsets> flattened: color_offsets[num_colors+1], element labels per colour in reference order. */
fb200_status fb200_colors_download(fb200_ctx* ctx, uint64_t* color_offsets, uint64_t* element_ids);
fb200_status fb200_colors_adopt(fb200_ctx* ctx, uint64_t num_colors, const uint64_t* color_offsets, const uint64_t* element_ids);

/* ---- assemble_into_csr (global.rs:133-182, 314-376) ---------------------------------------- */
/* u: global solution vector (solution_dim * num_nodes, host) or NULL (= zeros, as the reference's linear call sites pass).
 *    LAPLACE / LINEAR_ELASTIC do not depend on u (laplace.rs:62, materials.rs:110) and take the tuned stiffness kernels; STVK assembles the
 *    tangent stiffness at u (elliptic.rs:361-439 with u_grad per point; contraction C = I (2 mu a.Eb + lambda tr(E) a.b) + mu Fb Fa^T +
 *    lambda Fa Fb^T + mu (a.b) F F^T, fenris-solid/src/materials.rs:417-437), NEO_HOOKEAN likewise (C = lambda (F^-T a)(F^-T b)^T -
 *    alpha (F^-T b)(F^-T a)^T + mu (a.b) I, alpha = -mu + lambda log J, materials.rs:291-318), with the ATOMIC or COLORED scatter.
 * accumulate != 0: values += contributions (assemble_into_csr);  == 0: values = contributions (assemble()).
 * The _device form only enqueues work on the ctx stream (values stay in HBM; call fb200_synchronize to
 * collect deferred errors).  The host form uploads `values` first when accumulating, waits, and copies
 * the result back into `values` [nnz]. */
fb200_status fb200_assemble_into_csr_device(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature,
                                            const double* u, int32_t scatter_mode, int32_t accumulate);
fb200_status fb200_assemble_into_csr(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature,
                                     const double* u, int32_t scatter_mode, int32_t accumulate, double* values);
/* The same with a rule PER ELEMENT: CompactQuadratureTable (rules + element_to_rule_map, quadrature_table.rs:312-439); a
 * GeneralQuadratureTable (:57-210) is the special case of one rule per element (the Python mirror merges identical rules).
 * rules[r].data follows the operator as above; element_rule[e] < num_rules for every element of the space.
 * scatter_mode: ATOMIC or COLORED.  Elements are grouped by rule; each group runs with its own uniform tables.  u as above (STVK /
 * NEO_HOOKEAN assemble the tangent stiffness at u with the rule and the Lame data of each element). */
fb200_status fb200_assemble_into_csr_table_device(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, int32_t scatter_mode, int32_t accumulate);
fb200_status fb200_values_device(fb200_ctx* ctx, double** device_ptr, uint64_t* nnz);
fb200_status fb200_values_download(fb200_ctx* ctx, double* values);
fb200_status fb200_values_upload(fb200_ctx* ctx, const double* values);
/* Element matrices only (ElementMatrixAssembler::assemble_element_matrix, local.rs:77-103): K_e for elements
 * [first, first+count), each (s n)^2 doubles column-major like nalgebra's DMatrix. Test/debug surface. */
fb200_status fb200_element_matrices(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature,
                                    uint64_t first, uint64_t count, double* out);
/* The same view with the state the assembler was built with (ElementEllipticAssemblerBuilder::with_u, elliptic.rs:241-297): for the
 * state-dependent operators (FB200_STVK, FB200_NEO_HOOKEAN; fenris-solid/src/materials.rs:232-469) K_e is the tangent stiffness at u
 * (num_nodes * d doubles, NULL = zeros; fb200_element_matrices is this call with u = NULL); linear operators ignore u. */
fb200_status fb200_element_matrices_u(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature, const double* u,
                                      uint64_t first, uint64_t count, double* out);

/* ---- mass matrix, source vector, global vector assembly (the rest of "assemble a linear system", examples/poisson2d.rs:33-86) --- */
/* ElementMassAssembler through CsrAssembler / CsrParAssembler (src/assembly/local/mass.rs:127-159, 218-286): M_IJ = I_s * sum_q w_q |det J_q|
 * rho_q phi_I phi_J into the CSR values of the current pattern (s = the pattern's solution_dim).  quadrature->data = density per point
 * (Density, mass.rs:23-31; [num_points]).  scatter_mode: ATOMIC or COLORED.  accumulate as in fb200_assemble_into_csr. */
fb200_status fb200_assemble_mass_into_csr_device(fb200_ctx* ctx, const fb200_quadrature* quadrature, int32_t scatter_mode, int32_t accumulate);
fb200_status fb200_assemble_mass_into_csr(fb200_ctx* ctx, const fb200_quadrature* quadrature, int32_t scatter_mode, int32_t accumulate, double* values);
/* VectorAssembler / VectorParAssembler::assemble_vector_into with an ElementSourceAssembler (global.rs:569-686, source.rs:217-278):
 * out[s I + i] (+)= sum over elements and points of w_q |det J_q| phi_I(xi_q) f_i(x_q).  The source function is a caller-side closure in the
 * reference (SourceFunction::evaluate); here the caller passes its VALUES at the quadrature points: source_values[q * s + i] shared by all
 * elements (per_element = 0, e.g. gravity) or source_values[(e * num_points + q) * s + i] (per_element = 1; physical points from
 * fb200_physical_quadrature_points).  out: solution_dim * num_nodes doubles on the host; accumulate != 0 adds to its contents.
 * quadrature->data is not used.  scatter_mode: ATOMIC or COLORED. */
fb200_status fb200_assemble_vector(fb200_ctx* ctx, const fb200_quadrature* quadrature, int32_t solution_dim, const double* source_values,
                                   int32_t per_element, int32_t scatter_mode, int32_t accumulate, double* out);
/* ElementEllipticAssembler as ElementVectorAssembler / ElementScalarAssembler (src/assembly/local/elliptic.rs:342-359, 440-605): with
 * grad u = J^-T sum_I grad_ref phi_I (x) u_I (compute_volume_u_grad, :25-59) per quadrature point,
 *   vector  out[s I + i] (+)= sum w |det J| (g^T grad phi_I)_i      g = grad u (Laplace) | g^T = P(grad u) (LinearElasticMaterial / StVKMaterial / NeoHookeanMaterial stress)
 *   scalar  *energy = sum over elements and points of w |det J| psi(grad u)    psi = |grad u|^2 / 2 | mu eps:eps + lambda tr(eps)^2 / 2 | mu E:E + lambda tr(E)^2 / 2 | mu tr(E) - mu log J + lambda (log J)^2 / 2
 * u: solution_dim * num_nodes doubles on the host.  Vector: VectorAssembler / VectorParAssembler semantics (ATOMIC | COLORED, accumulate);
 * scalar: assemble_scalar (global.rs:697-722).  Errors: FB200_ERR_SINGULAR_JACOBIAN with the element index. */
fb200_status fb200_assemble_elliptic_vector(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature, const double* u,
                                            int32_t scatter_mode, int32_t accumulate, double* out);
fb200_status fb200_assemble_elliptic_scalar(fb200_ctx* ctx, const fb200_operator* op, const fb200_quadrature* quadrature, const double* u,
                                            double* energy);
/* The same two with a quadrature rule per element (rules / element_rule as fb200_assemble_into_csr_table_device). */
fb200_status fb200_assemble_elliptic_vector_table(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, int32_t scatter_mode, int32_t accumulate,
                                                  double* out);
fb200_status fb200_assemble_elliptic_scalar_table(fb200_ctx* ctx, const fb200_operator* op, uint32_t num_rules, const fb200_quadrature* rules,
                                                  const uint32_t* element_rule, const double* u, double* energy);
/* x_q = element.map_reference_coords(xi_q) for every element and point (FiniteElement::map_reference_coords; sub-parametric elements map
 * through their embedded linear element, hexahedron.rs:328-330): out[(e * num_points + q) * d + c]. */
fb200_status fb200_physical_quadrature_points(fb200_ctx* ctx, const fb200_quadrature* quadrature, double* out);

/* apply_homogeneous_dirichlet_bc_csr (src/assembly/global.rs:379-451) on the device-resident CSR values: rows of the listed nodes' dofs get
 * scale on the diagonal and 0 elsewhere, the rows coupled to them lose their entries in those columns; scale = |first non-zero diagonal
 * entry| (1 if none), returned in *scale (may be NULL).  The right-hand side counterpart (global.rs:479-495) is a host loop over s*node+i. */
fb200_status fb200_apply_homogeneous_dirichlet_bc_csr(fb200_ctx* ctx, uint64_t num_dirichlet_nodes, const uint64_t* nodes, double* scale);

/* ---- the consumer of the assembled matrix, on the device-resident CSR (no D2H of the matrix) ------------------------------- */
/* y = A x (host vectors of solution_dim * num_nodes doubles). */
fb200_status fb200_spmv(fb200_ctx* ctx, const double* x, double* y);
/* ConjugateGradient::solve_with_guess (fenris-sparse/src/cg.rs:364-480) with RelativeResidualCriterion(rel_tol) (cg.rs:85-124):
 * x holds the initial guess on entry and the solution on return; jacobi != 0 preconditions with the inverse diagonal, else identity;
 * max_iter = 0: unbounded.  Returns FB200_OK, FB200_ERR_NOT_CONVERGED (x = last iterate) or FB200_ERR_INDEFINITE.
 * *iterations = number of updates of x, *rel_residual = ||r|| / ||b|| (recursive residual) at exit. */
fb200_status fb200_cg_solve(fb200_ctx* ctx, const double* b, double* x, double rel_tol, uint64_t max_iter, int32_t jacobi, uint64_t* iterations,
                            double* rel_residual);

/* ---- multi-GPU: element partition + interface-row exchange --------------------------------- */
#define FB200_UNIQUE_ID_BYTES 128
fb200_status fb200_comm_unique_id(char id[FB200_UNIQUE_ID_BYTES]);
fb200_status fb200_comm_init(fb200_ctx* ctx, const char id[FB200_UNIQUE_ID_BYTES], int32_t rank, int32_t num_ranks);
/* Interface nodes = local nodes whose rows also receive contributions on other ranks.  packed_offsets[i] is the
 * position (in doubles) of node local_nodes[i]'s value block (solution_dim^2 * coupled-node-count doubles, identical on
 * every sharing rank because ghosts complete the pattern) inside a packed buffer of packed_len doubles that has the same
 * layout on every rank. */
fb200_status fb200_interface_set(fb200_ctx* ctx, uint64_t count, const uint64_t* local_nodes, const uint64_t* packed_offsets,
                                 uint64_t packed_len);
/* Optional, instead of fb200_interface_set: the interface as a list of PEERS.  Segment p = the local nodes shared with rank
 * peer_ranks[p], nodes[peer_begin[p] .. peer_begin[p+1]), in an order both ranks agree on (e.g. ascending global id); a node shared
 * with several ranks appears in each of their segments.  The exchange then is a neighbour exchange (ncclSend / ncclRecv of the packed
 * segments inside one group, sum on arrival) instead of a world all-reduce whose buffer grows with the number of ranks. */
fb200_status fb200_interface_set_peers(fb200_ctx* ctx, uint64_t num_peers, const int32_t* peer_ranks, const uint64_t* peer_begin,
                                       const uint64_t* nodes);
/* Sum the interface rows over the ranks that share them.  Peers set: pack per peer -> ncclSend/ncclRecv (one group) -> add the
 * received blocks; otherwise pack -> ncclAllReduce(sum, f64) -> unpack.  Over NVLink, enqueued on the ctx stream. */
fb200_status fb200_interface_allreduce(fb200_ctx* ctx);
/* Optional, after fb200_comm_init + fb200_interface_set_peers (COLLECTIVE over all ranks of the communicator: the ranks vote, and either
 * all of them enable it or none does): fuse the interface exchange into the
 * assembly kernel.  Every rank maps its neighbours' value arrays (CUDA IPC over NVLink) and learns where their copies of the shared rows
 * start; the Hex8 tile kernel's flush then adds this rank's partial sums of an interface row to the neighbour's copy as well
 * (red.global.add.f64 on the peer pointer), and fb200_interface_allreduce shrinks to a neighbour barrier - no pack, no ncclSend/ncclRecv,
 * no add pass.  Needs one neighbour per interface node and at most two neighbours per rank (slab-like partitions); otherwise
 * FB200_ERR_UNSUPPORTED and the packed exchange stays in use.  Assemblies that do not run the tile kernel keep the packed exchange too.
 * The mapping belongs to the current pattern: call again after a new pattern (on every rank). */
fb200_status fb200_interface_enable_p2p(fb200_ctx* ctx);

/* ---- host-side helpers restating the reference's generators (no GPU needed) ----------------- */
/* src/mesh/procedural.rs:216-277 / 286-403 / 46-93.  Call with vertices == NULL to query sizes. */
fb200_status fb200_gen_hex_mesh(uint64_t cells_x, uint64_t cells_y, uint64_t cells_z, double cell_size,
                                uint64_t* num_vertices, uint64_t* num_elements, double* vertices, uint64_t* connectivity);
fb200_status fb200_gen_tet_mesh(uint64_t cells_x, uint64_t cells_y, uint64_t cells_z, double cell_size,
                                uint64_t* num_vertices, uint64_t* num_elements, double* vertices, uint64_t* connectivity);
fb200_status fb200_gen_quad_mesh(uint64_t cells_x, uint64_t cells_y, double cell_size,
                                 uint64_t* num_vertices, uint64_t* num_elements, double* vertices, uint64_t* connectivity);
/* Hex27Mesh::from(&hex8_mesh) (src/mesh_convert.rs:85-166,227-330). vertices27 needs capacity for the returned count
 * (query with vertices27 == NULL). */
fb200_status fb200_hex27_from_hex8(uint64_t num_vertices, const double* vertices, uint64_t num_elements, const uint64_t* hex8,
                                   uint64_t* num_vertices27, double* vertices27, uint64_t* hex27);
/* Hex20Mesh::from(&hex8_mesh) (src/mesh_convert.rs:168-217, 481-488): same calling convention as fb200_hex27_from_hex8. */
fb200_status fb200_hex20_from_hex8(uint64_t num_vertices, const double* vertices, uint64_t num_elements, const uint64_t* hex8,
                                   uint64_t* num_vertices_out, double* vertices_out, uint64_t* hex20);
/* Tet10Mesh::from(&tet4_mesh) (src/mesh_convert.rs:42-83, 227-330, 444-452): the 4 vertices, then the midpoints of the edges (0,1) (1,2)
 * (0,2) (0,3) (2,3) (1,3), labelled in first-seen order.  Same calling convention as fb200_hex27_from_hex8. */
fb200_status fb200_tet10_from_tet4(uint64_t num_vertices, const double* vertices, uint64_t num_elements, const uint64_t* tet4,
                                   uint64_t* num_vertices_out, double* vertices_out, uint64_t* tet10);
/* Canonical stiffness quadrature of an element type (src/quadrature/canonical.rs:95,102-104,110-112):
 * query num_points with weights == NULL. points are point-major [num_points * dim]. */
fb200_status fb200_canonical_quadrature(int32_t element_type, int32_t* num_points, double* weights, double* points);
/* LameParameters::from(YoungPoisson) (fenris-solid/src/materials.rs:31-43). */
void fb200_lame_from_young_poisson(double young, double poisson, double* mu, double* lambda);

/* Host-only self check of the Hex8 tile lists (csrc/tiles.cpp) on a caller-supplied Hex8 mesh; no GPU needed.  Builds the node-block
 * map the way fb200_assemble_pattern does, the Morton order, the tile lists, and verifies them: every element scheduled exactly once,
 * rounds node-disjoint, every block (a, b) with u_a <= u_b mapped to the accumulator of its node pair, the flush list covering every
 * coupled pair of the tile exactly once with the right position in the CSR block row, complete flags, size limits.
 * stats[0..7] = tiles, max nodes, max accumulator positions, flush entries, complete nodes, bank-conflict share * 1e6, schedule positions,
 * max rounds.  Returns FB200_OK, FB200_ERR_UNSUPPORTED when the mesh cannot use tiles (repeated nodes), FB200_ERR_STATE + (*failed_check = id)
 * when a check fails.  The plain form builds the lists with first-writer ownership (fb200_set_tuning "hex8_owner_stores" = 1); the _ex form
 * takes the setting and also reports stats[8] = flush entries in the STORE segments, stats[9] = zero entries written by the owners of
 * shared rows.  Ownership checks: every row is stored completely by exactly one tile or listed for clearing (rows that ghost elements touch),
 * and a tile that reduces into a stored row waits for a lower-numbered tile. */
fb200_status fb200_tile_lists_selftest_ex(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                          uint64_t num_owned, int32_t owner_stores, uint64_t stats[10], int32_t* failed_check);
fb200_status fb200_tile_lists_selftest(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                       uint64_t num_owned, uint64_t stats[8], int32_t* failed_check);

/* Host-only self check of the Tet4 chunk lists (csrc/chunks.cpp; no GPU needed): every contribution (element, a, b) of every owned element
 * exactly once in the slot of its node block, contributors in ascending element order, destinations / row lengths consistent with the block
 * offsets, the "complete" flag exactly for rows whose node has all its elements - ghost elements of a partition included - inside the chunk,
 * the "interface" flag exactly for rows ghost elements touch.  stats[0..3] = chunks, slots, complete slots, interface slots. */
fb200_status fb200_chunk_lists_selftest(uint64_t num_nodes, const double* vertices, uint64_t num_elements, const uint64_t* connectivity,
                                        uint64_t num_owned, int32_t chunk_elems, int32_t solution_dim, uint64_t stats[4], int32_t* failed_check);

#ifdef __cplusplus
}
#endif
#endif /* FENRIS_B200_H */
