"""fenris_b200 - B200-native (sm_100a) global FEM operator assembly, a drop-in for fenris's
CsrAssembler / CsrParAssembler hot path.  See DESIGN.md, INTEGRATION.md and include/fenris_b200.h.

The numerical path is libfenris_b200.so (hand-written CUDA); there is no CPU fallback.
"""
from ._native import (ERR_COLORING, ERR_COLUMN_NOT_IN_PATTERN, ERR_CUDA, ERR_INDEFINITE, ERR_INDEX_OOB, ERR_NCCL,  # noqa: F401
                      ERR_NOT_CONVERGED, ERR_SHAPE,
                      ERR_SINGULAR_JACOBIAN, ERR_STATE, ERR_UNSUPPORTED, HEX8, HEX20, HEX27, LAPLACE, LINEAR_ELASTIC, NEO_HOOKEAN, STVK, OK, QUAD4,
                      SCATTER_ATOMIC, SCATTER_COLORED, SCATTER_GATHER, TET4, TET10, Fb200Error, SingularJacobianError)
from .api import (CompactQuadratureTable, CsrAssembler, CsrMatrix, CsrParAssembler, Density, GeneralQuadratureTable, DisjointSubsets, ElementConnectivityAssembler,  # noqa: F401
                  ElementEllipticAssembler, ElementEllipticAssemblerBuilder, ElementMassAssembler, ElementSourceAssembler, LameParameters, LaplaceOperator,
                  LinearElasticMaterial, MaterialEllipticOperator, Mesh, NeoHookeanMaterial, StVKMaterial, SparsityPattern, UniformQuadratureTable, VectorAssembler, VectorParAssembler, YoungPoisson,
                  apply_homogeneous_dirichlet_bc_csr, apply_homogeneous_dirichlet_bc_rhs, assemble_scalar, canonical_stiffness_quadrature, color_nodes, create_rectangular_uniform_hex_mesh,
                  create_rectangular_uniform_tet_mesh, create_unit_box_uniform_hex_mesh_3d, create_unit_box_uniform_tet_mesh_3d,
                  create_unit_square_uniform_quad_mesh_2d, hex20_mesh_from, hex27_mesh_from, tet10_mesh_from)
from .context import Context  # noqa: F401
