// XLA FFI custom-call handlers over the C-ABI of include/cmad_b200.h.
//
// The reference (sandialabs/cmad) is pure Python on JAX; its hot path is reached from JAX
// programs (cmad/fem/assembly.py:616-732 assemble_element_block; the per-point loops of
// cmad/objectives/mp_objective.py and cmad/cli/primal.py).  These handlers let that code hand
// its device arrays to libcmad_b200.so without leaving the XLA stream: every handler only
// fills the C-ABI structs with the buffers' device pointers and forwards the stream XLA
// gives it.  No allocation, no synchronisation, no host copy.
//
// Build (only where JAX's FFI headers exist - cmad_b200/xla/__init__.py:build()):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())")
//       -I<repo>/include -I$CUDA_HOME/include cmad_b200_xla.cc -L<repo>/cmad_b200/lib -lcmad_b200
// Registered from Python with jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(sym), "CUDA").
//
// Struct-valued attributes (material, newton settings) travel as uint8 arrays holding the
// bytes of cmadx_material_t / cmadx_newton_t (the Python side builds them with the ctypes
// mirrors of cmad_b200/_lib.py and checks cmadx_struct_sizes first).
#include <cstdint>
#include <cstring>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

#include "cmad_b200.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using Bytes = ffi::Span<const uint8_t>;

ffi::Error status(int rc) {
    if (rc == CMADX_OK) return ffi::Error::Success();
    if (rc == CMADX_EINVAL) return ffi::Error::InvalidArgument(cmadx_error_string(rc));
    if (rc == CMADX_EUNSUPPORTED) return ffi::Error(ffi::ErrorCode::kUnimplemented, cmadx_error_string(rc));
    if (rc == CMADX_ECUDA) return ffi::Error::Internal(cmadx_last_cuda_error());
    return ffi::Error::Internal(cmadx_error_string(rc));
}

template <class T>
bool unpack(Bytes bytes, T* out) {
    if (bytes.size() != sizeof(T)) return false;
    std::memcpy(out, bytes.data(), sizeof(T));
    return true;
}

#define CMADX_UNPACK(T, var, bytes)                                                         \
    T var;                                                                                  \
    if (!unpack(bytes, &var)) return ffi::Error::InvalidArgument(#bytes ": wrong struct size (ABI mismatch)")

// ------------------------------------------------------------------------------------
// Material-point update: replaces make_newton_solve(...)(xi_prev, params, U_ip, U_ip_prev) + its
// custom_jvp IFT rule (cmad/models/nonlinear_solver.py:88-174), newton_solve(model) (:14-85)
// and Model.cauchy / dC_dxi / dC_dxi_prev / dC_dp (cmad/models/model.py:121-166, 316-350)
// for a batch of points.  Layout: component-major [comps][n].
// ------------------------------------------------------------------------------------
ffi::Error MpUpdateImpl(cudaStream_t stream, F64 xi_prev, F64 strain, Bytes material, Bytes newton,
                        ffi::Span<const int32_t> active_pid, int32_t def_type, RF64 xi, RF64 sigma,
                        RF64 dsig_deps, RF64 dC_dp, RS32 iters, RS32 flags) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    CMADX_UNPACK(cmadx_newton_t, nw, newton);
    auto d = xi_prev.dimensions();
    auto e = strain.dimensions();
    if (d.size() != 2 || e.size() != 2 || d[1] != e[1]) return ffi::Error::InvalidArgument("xi_prev / strain: expected [comps][n]");
    cmadx_mp_buffers_t b;
    std::memset(&b, 0, sizeof b);
    b.n = d[1];
    b.ld = d[1];
    b.strain_comps = (int32_t)e[0];
    b.def_type = def_type;
    b.xi_prev = xi_prev.typed_data();
    b.strain = strain.typed_data();
    b.xi = xi->typed_data();
    b.sigma = sigma->typed_data();
    b.dsig_deps = dsig_deps->typed_data();
    b.dC_dp = active_pid.size() ? dC_dp->typed_data() : nullptr;
    b.iters = iters->typed_data();
    b.flags = flags->typed_data();
    return status(cmadx_mp_update(&mat, &nw, active_pid.data(), (int32_t)active_pid.size(), &b, stream));
}

// every per-point output of Model's AD products (cmad/models/model.py:179-189, 316-374)
ffi::Error MpUpdateFullImpl(cudaStream_t stream, F64 xi_prev, F64 strain, Bytes material, Bytes newton,
                            ffi::Span<const int32_t> active_pid, int32_t def_type, RF64 xi, RF64 sigma,
                            RF64 dsig_deps, RF64 dxi_deps, RF64 dC_dp, RF64 dC_dxi, RF64 dC_dxi_prev,
                            RF64 C, RF64 cnorm, RS32 iters, RS32 flags) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    CMADX_UNPACK(cmadx_newton_t, nw, newton);
    auto d = xi_prev.dimensions();
    auto e = strain.dimensions();
    if (d.size() != 2 || e.size() != 2 || d[1] != e[1]) return ffi::Error::InvalidArgument("xi_prev / strain: expected [comps][n]");
    cmadx_mp_buffers_t b;
    std::memset(&b, 0, sizeof b);
    b.n = d[1];
    b.ld = d[1];
    b.strain_comps = (int32_t)e[0];
    b.def_type = def_type;
    b.xi_prev = xi_prev.typed_data();
    b.strain = strain.typed_data();
    b.xi = xi->typed_data();
    b.sigma = sigma->typed_data();
    b.dsig_deps = dsig_deps->typed_data();
    b.dxi_deps = dxi_deps->typed_data();
    b.dC_dp = active_pid.size() ? dC_dp->typed_data() : nullptr;
    b.dC_dxi = dC_dxi->typed_data();
    b.dC_dxi_prev = dC_dxi_prev->typed_data();
    b.C = C->typed_data();
    b.cnorm = cnorm->typed_data();
    b.iters = iters->typed_data();
    b.flags = flags->typed_data();
    return status(cmadx_mp_update(&mat, &nw, active_pid.data(), (int32_t)active_pid.size(), &b, stream));
}

// Model._jacobian[DU], Model.dcauchy[DXI | DPARAMS] and jacfwd(cauchy, DU) at given states
// (cmad/models/model.py:121-160), component-major [comps][n]
ffi::Error MpModelPartialsImpl(cudaStream_t stream, F64 xi, F64 xi_prev, F64 strain, Bytes material,
                               ffi::Span<const int32_t> active_pid, RF64 dC_deps, RF64 dsig_dxi, RF64 dsig_deps,
                               RF64 dsig_dp) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    auto d = xi.dimensions();
    auto dp = xi_prev.dimensions();
    auto e = strain.dimensions();
    if (d.size() != 2 || dp.size() != 2 || e.size() != 2 || d[0] != 7 || dp[0] != 7 || d[1] != e[1] || dp[1] != d[1])
        return ffi::Error::InvalidArgument("xi / xi_prev / strain: expected [7][n], [7][n], [6|9][n]");
    cmadx_mp_partials_t p;
    std::memset(&p, 0, sizeof p);
    p.n = d[1];
    p.ld = d[1];
    p.strain_comps = (int32_t)e[0];
    p.xi = xi.typed_data();
    p.xi_prev = xi_prev.typed_data();
    p.strain = strain.typed_data();
    p.dC_deps = dC_deps->typed_data();
    p.dsig_dxi = dsig_dxi->typed_data();
    p.dsig_deps = dsig_deps->typed_data();
    p.dsig_dp = active_pid.size() ? dsig_dp->typed_data() : nullptr;
    return status(cmadx_mp_model_partials(&mat, active_pid.data(), (int32_t)active_pid.size(), &p, stream));
}

// sorted_eigen_decomposition (cmad/util/jax_eigen_decomposition.py:86-171) over a batch: A6 [6][n] ->
// eigenvalues [3][n] ascending, eigenvectors [9][n] (row 3 m + k = component m of vector k)
ffi::Error Sym3EighImpl(cudaStream_t stream, F64 A6, RF64 w, RF64 V) {
    auto d = A6.dimensions();
    if (d.size() != 2 || d[0] != 6) return ffi::Error::InvalidArgument("A6: expected [6][n]");
    return status(cmadx_sym3_eigh(d[1], d[1], A6.typed_data(), w->typed_data(), V->typed_data(), stream));
}

// ------------------------------------------------------------------------------------
// FE element block.  The arrays are the reference's own, unchanged: u_gather_eq (int32),
// grad_N_phys (n_e, n_ip, n_b, 3), iso_jac_det (n_e, n_ip), quad_w, xi (n_e, n_ip, n_xi)
// (cmad/fem/kernel_arrays.py:58-228, cmad/fem/precompute.py:58-122).
// ------------------------------------------------------------------------------------
ffi::Error fill_block(cmadx_fe_block_t* b, const S32& elem_eq, const F64& U, const F64& xi_prev,
                      const F64& grad_N, const F64& det, const F64& quad_w) {
    std::memset(b, 0, sizeof *b);
    auto g = grad_N.dimensions();                    // (n_e, n_ip, n_b, 3)
    if (g.size() != 4 || g[3] != 3) return ffi::Error::InvalidArgument("grad_N: expected (n_e, n_ip, n_b, 3)");
    b->n_elems = g[0];
    b->n_ip = (int32_t)g[1];
    b->n_basis = (int32_t)g[2];
    b->n_dofs = (int64_t)U.element_count();
    if ((int64_t)elem_eq.element_count() != g[0] * g[2] * 3) return ffi::Error::InvalidArgument("elem_eq: expected (n_e, n_b * 3)");
    if ((int64_t)det.element_count() != g[0] * g[1] || (int64_t)quad_w.element_count() != g[1])
        return ffi::Error::InvalidArgument("det / quad_w: shapes do not match grad_N");
    if ((int64_t)xi_prev.element_count() != g[0] * g[1] * 7) return ffi::Error::InvalidArgument("xi_prev: expected (n_e, n_ip, 7)");
    b->elem_eq = elem_eq.typed_data();
    b->U = U.typed_data();
    b->xi_prev = xi_prev.typed_data();
    b->grad_N = grad_N.typed_data();
    b->det = det.typed_data();
    b->quad_w = quad_w.typed_data();
    return ffi::Error::Success();
}

// assemble_element_block, COUPLED mode (cmad/fem/assembly.py:616-732): R_elem, the COO `vals`
// stream (elem, row dof, col dof), xi_solved
ffi::Error FeBlockImpl(cudaStream_t stream, S32 elem_eq, F64 U, F64 xi_prev, F64 grad_N, F64 det, F64 quad_w,
                       Bytes material, Bytes newton, RF64 R_elem, RF64 K_elem, RF64 xi) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    CMADX_UNPACK(cmadx_newton_t, nw, newton);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    b.R_elem = R_elem->typed_data();
    b.K_elem = K_elem->typed_data();
    b.xi = xi->typed_data();
    return status(cmadx_fe_block_assemble(&mat, &nw, &b, stream));
}

// assemble_element_block_residual (cmad/fem/assembly.py:735-813; per_element_R_coupled :538-613)
ffi::Error FeBlockResidualImpl(cudaStream_t stream, S32 elem_eq, F64 U, F64 xi_prev, F64 grad_N, F64 det,
                               F64 quad_w, Bytes material, Bytes newton, RF64 R_elem, RF64 xi) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    CMADX_UNPACK(cmadx_newton_t, nw, newton);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    b.R_elem = R_elem->typed_data();
    b.xi = xi->typed_data();
    return status(cmadx_fe_block_assemble(&mat, &nw, &b, stream));
}

ffi::Error fill_mixed(cmadx_fe_mixed_t* m, const cmadx_fe_block_t& b, const S32& elem_eq_p, const F64& N,
                      const F64& h, double stab_mult) {
    std::memset(m, 0, sizeof *m);
    if ((int64_t)elem_eq_p.element_count() != b.n_elems * b.n_basis) return ffi::Error::InvalidArgument("elem_eq_p: expected (n_e, n_b)");
    if ((int64_t)N.element_count() != (int64_t)b.n_ip * b.n_basis) return ffi::Error::InvalidArgument("N: expected (n_ip, n_b)");
    if ((int64_t)h.element_count() != b.n_elems) return ffi::Error::InvalidArgument("h: expected (n_e,)");
    m->elem_eq_p = elem_eq_p.typed_data();
    m->N = N.typed_data();
    m->h = h.typed_data();
    m->stab_mult = stab_mult;
    return ffi::Error::Success();
}

// the same for SmallDispEquilibrium(mixed=True) (cmad/global_residuals/small_disp_equilibrium.py:87-111):
// two residual blocks and the four (r, s)-ordered COO streams (assembly.py:722-732)
ffi::Error FeBlockMixedImpl(cudaStream_t stream, S32 elem_eq, S32 elem_eq_p, F64 U, F64 xi_prev, F64 grad_N,
                            F64 det, F64 quad_w, F64 N, F64 h, Bytes material, Bytes newton, double stab_mult,
                            RF64 R_u, RF64 R_p, RF64 K_uu, RF64 K_up, RF64 K_pu, RF64 K_pp, RF64 xi) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    CMADX_UNPACK(cmadx_newton_t, nw, newton);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    cmadx_fe_mixed_t m;
    if (ffi::Error e = fill_mixed(&m, b, elem_eq_p, N, h, stab_mult); e.failure()) return e;
    b.R_elem = R_u->typed_data();
    b.K_elem = K_uu->typed_data();
    b.xi = xi->typed_data();
    m.R_p_elem = R_p->typed_data();
    m.K_up = K_up->typed_data();
    m.K_pu = K_pu->typed_data();
    m.K_pp = K_pp->typed_data();
    return status(cmadx_fe_block_assemble_mixed(&mat, &nw, &b, &m, stream));
}

// ------------------------------------------------------------------------------------
// K6: the derivative rules the reference obtains by jax.jvp / transposition through the FE
// Newton's IFT rule (cmad/fem/nonlinear_solver.py:490-537, cmad/models/nonlinear_solver.py:158-171)
// ------------------------------------------------------------------------------------
// forward: (dp, dxi_prev, dU) -> (dR_elem, dxi) at the converged state xi_state
ffi::Error FeBlockJvpImpl(cudaStream_t stream, S32 elem_eq, F64 U, F64 xi_prev, F64 grad_N, F64 det, F64 quad_w,
                          F64 xi_state, F64 dxi_prev, F64 dU, Bytes material, ffi::Span<const int32_t> active_pid,
                          ffi::Span<const double> dp, RF64 dR_elem, RF64 dxi) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    if (dp.size() != active_pid.size()) return ffi::Error::InvalidArgument("dp: one entry per active parameter");
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    if (xi_state.element_count() != xi_prev.element_count() || dxi_prev.element_count() != xi_prev.element_count() ||
        dU.element_count() != U.element_count())
        return ffi::Error::InvalidArgument("xi_state / dxi_prev / dU: shapes do not match xi_prev / U");
    b.R_elem = dR_elem->typed_data();
    b.xi = dxi->typed_data();
    return status(cmadx_fe_block_jvp(&mat, active_pid.data(), (int32_t)active_pid.size(), dp.data(), &b,
                                     xi_state.typed_data(), dxi_prev.typed_data(), dU.typed_data(), stream));
}

// reverse: (Rbar, xibar) -> (pbar, xibar_prev); `workspace` is a scratch result buffer of
// cmadx_fe_vjp_workspace_bytes(n_e, n_ip, n_active) / 8 doubles (the Python side sizes it)
ffi::Error FeBlockVjpImpl(cudaStream_t stream, S32 elem_eq, F64 U, F64 xi_prev, F64 grad_N, F64 det, F64 quad_w,
                          F64 xi_state, F64 Rbar, F64 xibar, Bytes material, ffi::Span<const int32_t> active_pid,
                          RF64 pbar, RF64 xibar_prev, RF64 workspace) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    const int32_t na = (int32_t)active_pid.size();
    if ((int64_t)pbar->element_count() != na) return ffi::Error::InvalidArgument("pbar: one entry per active parameter");
    if ((int64_t)workspace->element_count() * 8 < cmadx_fe_vjp_workspace_bytes(b.n_elems, b.n_ip, na))
        return ffi::Error::InvalidArgument("workspace: smaller than cmadx_fe_vjp_workspace_bytes");
    if (Rbar.element_count() != U.element_count() || xibar.element_count() != xi_prev.element_count())
        return ffi::Error::InvalidArgument("Rbar / xibar: shapes do not match U / xi_prev");
    b.xi = xibar_prev->typed_data();
    return status(cmadx_fe_block_vjp(&mat, active_pid.data(), na, &b, xi_state.typed_data(), Rbar.typed_data(),
                                     xibar.typed_data(), pbar->typed_data(), workspace->typed_data(), stream));
}

// displacement cotangent of the converged block: per-point rows (n_e * n_ip, 3 n_b)
ffi::Error FeBlockVjpDispImpl(cudaStream_t stream, S32 elem_eq, F64 U, F64 xi_prev, F64 grad_N, F64 det, F64 quad_w,
                              F64 xi_state, F64 Rbar, F64 xibar, Bytes material, RF64 Ubar_ip, RF64 xibar_prev) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    if ((int64_t)Ubar_ip->element_count() != b.n_elems * b.n_ip * b.n_basis * 3)
        return ffi::Error::InvalidArgument("Ubar_ip: expected (n_e * n_ip, 3 n_b)");
    b.xi = xibar_prev->typed_data();
    return status(cmadx_fe_block_vjp_disp(&mat, &b, nullptr, xi_state.typed_data(), Rbar.typed_data(),
                                          xibar.typed_data(), Ubar_ip->typed_data(), stream));
}

// mixed u-p JVP / VJP over both residual blocks
ffi::Error FeBlockJvpMixedImpl(cudaStream_t stream, S32 elem_eq, S32 elem_eq_p, F64 U, F64 xi_prev, F64 grad_N,
                               F64 det, F64 quad_w, F64 N, F64 h, F64 xi_state, F64 dxi_prev, F64 dU, Bytes material,
                               ffi::Span<const int32_t> active_pid, ffi::Span<const double> dp, double stab_mult,
                               RF64 dR_u, RF64 dR_p, RF64 dxi) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    if (dp.size() != active_pid.size()) return ffi::Error::InvalidArgument("dp: one entry per active parameter");
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    cmadx_fe_mixed_t m;
    if (ffi::Error e = fill_mixed(&m, b, elem_eq_p, N, h, stab_mult); e.failure()) return e;
    b.R_elem = dR_u->typed_data();
    b.xi = dxi->typed_data();
    m.R_p_elem = dR_p->typed_data();
    return status(cmadx_fe_block_jvp_mixed(&mat, active_pid.data(), (int32_t)active_pid.size(), dp.data(), &b, &m,
                                           xi_state.typed_data(), dxi_prev.typed_data(), dU.typed_data(), stream));
}

ffi::Error FeBlockVjpMixedImpl(cudaStream_t stream, S32 elem_eq, S32 elem_eq_p, F64 U, F64 xi_prev, F64 grad_N,
                               F64 det, F64 quad_w, F64 N, F64 h, F64 xi_state, F64 Rbar, F64 xibar, Bytes material,
                               ffi::Span<const int32_t> active_pid, double stab_mult, RF64 pbar, RF64 xibar_prev,
                               RF64 workspace) {
    CMADX_UNPACK(cmadx_material_t, mat, material);
    cmadx_fe_block_t b;
    if (ffi::Error e = fill_block(&b, elem_eq, U, xi_prev, grad_N, det, quad_w); e.failure()) return e;
    cmadx_fe_mixed_t m;
    if (ffi::Error e = fill_mixed(&m, b, elem_eq_p, N, h, stab_mult); e.failure()) return e;
    const int32_t na = (int32_t)active_pid.size();
    if ((int64_t)workspace->element_count() * 8 < cmadx_fe_vjp_workspace_bytes(b.n_elems, b.n_ip, na))
        return ffi::Error::InvalidArgument("workspace: smaller than cmadx_fe_vjp_workspace_bytes");
    b.xi = xibar_prev->typed_data();
    return status(cmadx_fe_block_vjp_mixed(&mat, active_pid.data(), na, &b, &m, xi_state.typed_data(),
                                           Rbar.typed_data(), xibar.typed_data(), pbar->typed_data(),
                                           workspace->typed_data(), stream));
}

}  // namespace

#define CMADX_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxMpUpdate, MpUpdateImpl,
    CMADX_STREAM.Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<Bytes>("newton").Attr<ffi::Span<const int32_t>>("active_pid").Attr<int32_t>("def_type")
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<S32>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxMpUpdateFull, MpUpdateFullImpl,
    CMADX_STREAM.Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<Bytes>("newton").Attr<ffi::Span<const int32_t>>("active_pid").Attr<int32_t>("def_type")
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
        .Ret<S32>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxMpModelPartials, MpModelPartialsImpl,
    CMADX_STREAM.Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<ffi::Span<const int32_t>>("active_pid")
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxSym3Eigh, Sym3EighImpl,
    CMADX_STREAM.Arg<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlock, FeBlockImpl,
    CMADX_STREAM.Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<Bytes>("newton")
        .Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockResidual, FeBlockResidualImpl,
    CMADX_STREAM.Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<Bytes>("newton")
        .Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockMixed, FeBlockMixedImpl,
    CMADX_STREAM.Arg<S32>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<Bytes>("newton").Attr<double>("stab_mult")
        .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockJvp, FeBlockJvpImpl,
    CMADX_STREAM.Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<ffi::Span<const int32_t>>("active_pid").Attr<ffi::Span<const double>>("dp")
        .Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockVjp, FeBlockVjpImpl,
    CMADX_STREAM.Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<ffi::Span<const int32_t>>("active_pid")
        .Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockVjpDisp, FeBlockVjpDispImpl,
    CMADX_STREAM.Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material")
        .Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockJvpMixed, FeBlockJvpMixedImpl,
    CMADX_STREAM.Arg<S32>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<ffi::Span<const int32_t>>("active_pid").Attr<ffi::Span<const double>>("dp")
        .Attr<double>("stab_mult")
        .Ret<F64>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(CmadxFeBlockVjpMixed, FeBlockVjpMixedImpl,
    CMADX_STREAM.Arg<S32>().Arg<S32>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
        .Arg<F64>().Arg<F64>().Arg<F64>()
        .Attr<Bytes>("material").Attr<ffi::Span<const int32_t>>("active_pid").Attr<double>("stab_mult")
        .Ret<F64>().Ret<F64>().Ret<F64>());
