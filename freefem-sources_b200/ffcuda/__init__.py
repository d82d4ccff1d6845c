"""ctypes harness over the C ABI of libffcuda_core.so (include/ffcuda.h).

This is NOT the product's host side (that is the C++ FreeFEM plugin, plugin/ffcuda.cpp): it only lets tests/ and
bench.py drive exactly the entry points the plugin calls.  No numpy/torch arithmetic happens here; every compute
call goes to the CUDA library and fails loudly (FfcudaError) when the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libffcuda_core.so")
OP_ID, OP_DX, OP_DY, OP_DZ = 0, 1, 2, 6

# every symbol include/ffcuda.h declares (checked against the header by tests/test_abi.py)
SYMBOLS = """ffcuda_ctx_create ffcuda_ctx_destroy ffcuda_last_error ffcuda_ctx_sync ffcuda_ctx_set_stream ffcuda_ctx_get_stream ffcuda_ctx_set_option
ffcuda_prof_enable ffcuda_prof_reset ffcuda_prof_get ffcuda_launch_count ffcuda_mesh_upload ffcuda_mesh_cube ffcuda_mesh_square ffcuda_mesh_buildlayers
ffcuda_mesh_info ffcuda_mesh_download ffcuda_mesh_destroy ffcuda_space_create ffcuda_space_info ffcuda_space_download_dofs
ffcuda_space_destroy ffcuda_partition_rows_local ffcuda_matrix_from_csr_distributed ffcuda_space_create_distributed ffcuda_partition_local_nodes ffcuda_symbolic ffcuda_pattern_info ffcuda_pattern_download ffcuda_pattern_download_async ffcuda_pattern_lower_nnz ffcuda_pattern_download_lower
ffcuda_matrix_download_lower ffcuda_matrix_from_csr_lower ffcuda_pattern_destroy ffcuda_matrix_create
ffcuda_matrix_from_csr ffcuda_matrix_info ffcuda_matrix_download ffcuda_matrix_upload ffcuda_matrix_destroy ffcuda_vec_create
ffcuda_vec_upload ffcuda_vec_download ffcuda_vec_fill ffcuda_vec_ptr ffcuda_vec_destroy ffcuda_assemble_bilinear ffcuda_assemble_bilinear_qcoef
ffcuda_assemble_linear ffcuda_assemble_linear_qvalues ffcuda_assemble_linear_qterms ffcuda_assemble_linear_boundary ffcuda_assemble_bilinear_boundary ffcuda_assemble_linear_boundary_qvalues ffcuda_assemble_bilinear_boundary_qcoef ffcuda_bc_from_pairs ffcuda_bc_from_labels ffcuda_bc_count ffcuda_matrix_apply_bc ffcuda_vec_apply_bc
ffcuda_vec_set_bc_values ffcuda_bc_destroy ffcuda_spmv ffcuda_cg ffcuda_cg_host ffcuda_gmres ffcuda_gmres_host ffcuda_comm_unique_id ffcuda_comm_init
ffcuda_comm_finalize ffcuda_mesh_cube_distributed ffcuda_mesh_local_to_global ffcuda_quadrature ffcuda_partition_cube
ffcuda_partition_rcb ffcuda_partition_local ffcuda_matrix_export_device ffcuda_matrix_download_coo ffcuda_matrix_write_morse ffcuda_mesh_adjacency ffcuda_mesh_upload_distributed ffcuda_cg_stop_threshold ffcuda_fe_table
ffcuda_assemble_bilinear_rect ffcuda_matrix_shape ffcuda_matrix_download_csr""".split()


class FfcudaError(RuntimeError):
    pass


PART_FIELDS = ("L0 nown c_lo ncl nv_owned nv_local nt_local nbr_lo nbr_hi send_off_lo send_off_hi recv_off_lo recv_off_hi "
               "layer has_lower has_upper").split()


def partition_cube(nx, ny, nz, rank, nranks):
    """slab partition of cube(nx,ny,nz) for `rank` of `nranks` (host arithmetic only)."""
    out = (C.c_int64 * 16)()
    _ck(lib().ffcuda_partition_cube(nx, ny, nz, rank, nranks, out))
    return dict(zip(PART_FIELDS, [int(v) for v in out]))


def partition_rcb(xyz, nparts):
    """recursive coordinate bisection of the vertices (host arithmetic only): part[v] in [0, nparts)"""
    xyz = _f64(xyz)
    part = np.zeros(xyz.shape[0], np.int32)
    _ck(lib().ffcuda_partition_rcb(xyz.shape[1], xyz.shape[0], _p(xyz), int(nparts), _p(part)))
    return part


def partition_local(dim, nv, conn, part, rank, nranks):
    """the local problem of `rank` for a vertex partition (host arithmetic only): owned + ghost vertices, local elements,
    neighbours with their contiguous receive ranges and the gather lists to send"""
    conn, part = _i32(conn), _i32(part)
    sz = (C.c_int64 * 8)()
    args = (dim, nv, conn.shape[0], _p(conn), _p(part), rank, nranks, sz)
    _ck(lib().ffcuda_partition_local(*args, None, None, None, None, None, None, None))
    no, ng, ne, nn, ns = (int(sz[i]) for i in range(5))
    out = dict(nowned=no, l2g=np.zeros(no + ng, np.int32), elems=np.zeros(ne, np.int32), nbr=np.zeros(nn, np.int32),
               recv_off=np.zeros(nn, np.int32), recv_cnt=np.zeros(nn, np.int32), send_ptr=np.zeros(nn + 1, np.int32),
               send_idx=np.zeros(ns, np.int32))
    _ck(lib().ffcuda_partition_local(*args, _p(out["l2g"]), _p(out["elems"]), _p(out["nbr"]), _p(out["recv_off"]), _p(out["recv_cnt"]),
                                     _p(out["send_ptr"]), _p(out["send_idx"])))
    return out


def partition_local_nodes(elem2node, nnodes, part, rank, nranks):
    """the same for any element -> node table and node partition (P2 spaces on a distributed mesh)"""
    e2n, part = _i32(elem2node), _i32(part)
    sz = (C.c_int64 * 8)()
    args = (e2n.shape[1], int(nnodes), e2n.shape[0], _p(e2n), _p(part), rank, nranks, sz)
    _ck(lib().ffcuda_partition_local_nodes(*args, None, None, None, None, None, None, None))
    no, ng, ne, nn, ns = (int(sz[i]) for i in range(5))
    out = dict(nowned=no, l2g=np.zeros(no + ng, np.int32), elems=np.zeros(ne, np.int32), nbr=np.zeros(nn, np.int32),
               recv_off=np.zeros(nn, np.int32), recv_cnt=np.zeros(nn, np.int32), send_ptr=np.zeros(nn + 1, np.int32),
               send_idx=np.zeros(ns, np.int32))
    _ck(lib().ffcuda_partition_local_nodes(*args, _p(out["l2g"]), _p(out["elems"]), _p(out["nbr"]), _p(out["recv_off"]),
                                           _p(out["recv_cnt"]), _p(out["send_ptr"]), _p(out["send_idx"])))
    return out


def partition_rows_local(n, rowptr, colind, rank, nranks):
    """the local problem of `rank` when a host CSR matrix is shared out by contiguous row blocks (host arithmetic only)"""
    rowptr, colind = _i32(rowptr), _i32(colind)
    sz = (C.c_int64 * 8)()
    args = (int(n), _p(rowptr), _p(colind), rank, nranks, sz)
    _ck(lib().ffcuda_partition_rows_local(*args, None, None, None, None, None, None, None, None))
    no, ng, lnnz, nn, ns, lo = (int(sz[i]) for i in range(6))
    out = dict(nowned=no, first=lo, l2g=np.zeros(no + ng, np.int32), rowptr=np.zeros(no + 1, np.int32), colind=np.zeros(lnnz, np.int32),
               nbr=np.zeros(nn, np.int32), recv_off=np.zeros(nn, np.int32), recv_cnt=np.zeros(nn, np.int32),
               send_ptr=np.zeros(nn + 1, np.int32), send_idx=np.zeros(ns, np.int32))
    _ck(lib().ffcuda_partition_rows_local(*args, _p(out["l2g"]), _p(out["rowptr"]), _p(out["colind"]), _p(out["nbr"]), _p(out["recv_off"]),
                                          _p(out["recv_cnt"]), _p(out["send_ptr"]), _p(out["send_idx"])))
    return out


def quadrature(dim, qforder=6):
    """FreeFEM's default rule for int2d/int3d(Th, qforder=...): (points nq x dim, weights nq)."""
    n = C.c_int()
    _ck(lib().ffcuda_quadrature(dim, qforder, C.byref(n), None, None))
    pts, w = np.zeros((n.value, dim)), np.zeros(n.value)
    _ck(lib().ffcuda_quadrature(dim, qforder, C.byref(n), _p(pts), _p(w)))
    return pts, w


class BTerm(C.Structure):
    _fields_ = [("ucomp", C.c_int32), ("uop", C.c_int32), ("vcomp", C.c_int32), ("vop", C.c_int32), ("coef", C.c_double)]


class LTerm(C.Structure):
    _fields_ = [("vcomp", C.c_int32), ("vop", C.c_int32), ("coef", C.c_double)]


_lib = None


def lib():
    """Load the CUDA library; there is no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FfcudaError(f"{LIB_PATH} is missing: build it with `make -C freefem-sources_b200` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.ffcuda_last_error.restype = C.c_char_p
        L.ffcuda_last_error.argtypes = [C.c_void_p]
        L.ffcuda_launch_count.restype = C.c_int64
        L.ffcuda_launch_count.argtypes = [C.c_void_p]
        L.ffcuda_vec_ptr.restype = C.c_void_p
        L.ffcuda_vec_ptr.argtypes = [C.c_void_p]
        L.ffcuda_ctx_get_stream.restype = C.c_void_p
        L.ffcuda_ctx_get_stream.argtypes = [C.c_void_p]
        for name in SYMBOLS:
            f = getattr(L, name)
            if name.endswith("_destroy"):
                f.restype = None
                f.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ck(rc, ctx=None):
    if rc != 0:
        msg = lib().ffcuda_last_error(ctx)
        if not msg:
            msg = lib().ffcuda_last_error(None)
        raise FfcudaError((msg or b"unknown error").decode())


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _table(t):
    """a q-table argument: a numpy array (host) or a device Vec (its pointer is handed over: the table is used where it lies)"""
    if isinstance(t, Vec):
        return t, C.c_void_p(t.ptr())
    t = _f64(t)
    return t, _p(t)


def _h(obj):
    return C.c_void_p(obj.h)


class _Handle:
    _destroy = None

    def __init__(self, h, ctx):
        self.h = h
        self.ctx = ctx

    def close(self):
        if getattr(self, "h", None):
            getattr(lib(), self._destroy)(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context(_Handle):
    _destroy = "ffcuda_ctx_destroy"

    def __init__(self, device=0):
        out = C.c_void_p()
        _ck(lib().ffcuda_ctx_create(int(device), C.byref(out)))
        super().__init__(out.value, self)

    def sync(self):
        _ck(lib().ffcuda_ctx_sync(_h(self)), self.h)

    def set_stream(self, cuda_stream):
        _ck(lib().ffcuda_ctx_set_stream(_h(self), C.c_void_p(cuda_stream)), self.h)

    def set_option(self, name, value):
        _ck(lib().ffcuda_ctx_set_option(_h(self), name.encode(), int(value)), self.h)

    def prof_enable(self, on=True):
        _ck(lib().ffcuda_prof_enable(_h(self), int(on)), self.h)

    def prof_reset(self):
        _ck(lib().ffcuda_prof_reset(_h(self)), self.h)

    def prof_get(self, prefix=""):
        ms, n = C.c_double(), C.c_int64()
        _ck(lib().ffcuda_prof_get(_h(self), prefix.encode(), C.byref(ms), C.byref(n)), self.h)
        return ms.value, n.value

    def launch_count(self):
        return lib().ffcuda_launch_count(_h(self))

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        _ck(lib().ffcuda_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, nranks, id128):
        _ck(lib().ffcuda_comm_init(_h(self), rank, nranks, id128), self.h)

    def comm_finalize(self):
        _ck(lib().ffcuda_comm_finalize(_h(self)), self.h)

    # ---- factories
    def mesh_upload(self, dim, xyz, conn, elab=None, bconn=None, blab=None, belem=None, bface=None):
        xyz, conn, elab = _f64(xyz), _i32(conn), _i32(elab)
        bconn, blab, belem, bface = _i32(bconn), _i32(blab), _i32(belem), _i32(bface)
        nbe = 0 if blab is None else len(blab)
        out = C.c_void_p()
        _ck(lib().ffcuda_mesh_upload(_h(self), dim, xyz.shape[0], _p(xyz), conn.shape[0], _p(conn), _p(elab), nbe, _p(bconn),
                                     _p(blab), _p(belem), _p(bface), C.byref(out)), self.h)
        return Mesh(out.value, self)

    def mesh_upload_distributed(self, dim, nv_owned, xyz, conn, elab, bconn, blab, belem, bface, gid, nbr, recv_off, recv_cnt,
                                send_ptr, send_idx):
        """the local problem of this rank for any vertex partition (arrays in LOCAL numbering, see include/ffcuda.h)"""
        xyz, conn, elab = _f64(xyz), _i32(conn), _i32(elab)
        bconn, blab, belem, bface = _i32(bconn), _i32(blab), _i32(belem), _i32(bface)
        gid = np.ascontiguousarray(gid, dtype=np.int64)
        nbr, recv_off, recv_cnt, send_ptr, send_idx = _i32(nbr), _i32(recv_off), _i32(recv_cnt), _i32(send_ptr), _i32(send_idx)
        nbe = 0 if blab is None else len(blab)
        out = C.c_void_p()
        _ck(lib().ffcuda_mesh_upload_distributed(_h(self), dim, int(nv_owned), xyz.shape[0], _p(xyz), conn.shape[0], _p(conn), _p(elab),
                                                 nbe, _p(bconn), _p(blab), _p(belem), _p(bface), _p(gid), len(nbr), _p(nbr),
                                                 _p(recv_off), _p(recv_cnt), _p(send_ptr), _p(send_idx), C.byref(out)), self.h)
        return Mesh(out.value, self)

    def mesh_cube(self, nx, ny, nz, distributed=False):
        out = C.c_void_p()
        f = lib().ffcuda_mesh_cube_distributed if distributed else lib().ffcuda_mesh_cube
        _ck(f(_h(self), nx, ny, nz, C.byref(out)), self.h)
        return Mesh(out.value, self)

    def mesh_square(self, nx, ny):
        out = C.c_void_p()
        _ck(lib().ffcuda_mesh_square(_h(self), nx, ny, C.byref(out)), self.h)
        return Mesh(out.value, self)

    def vec(self, n):
        out = C.c_void_p()
        _ck(lib().ffcuda_vec_create(_h(self), int(n), C.byref(out)), self.h)
        return Vec(out.value, self, int(n))

    def vec_from(self, host):
        host = _f64(host)
        v = self.vec(len(host))
        v.upload(host)
        return v

    def matrix_from_csr_distributed(self, L, vals):
        """rows of this rank from the local problem L (partition_rows_local) and the values of its rows"""
        vals = _f64(vals)
        out = C.c_void_p()
        _ck(lib().ffcuda_matrix_from_csr_distributed(_h(self), L["nowned"], len(L["l2g"]), C.c_int64(len(L["colind"])), _p(L["rowptr"]),
                                                     _p(L["colind"]), _p(vals), len(L["nbr"]), _p(L["nbr"]), _p(L["recv_off"]),
                                                     _p(L["recv_cnt"]), _p(L["send_ptr"]), _p(L["send_idx"]), C.byref(out)), self.h)
        return Matrix(out.value, self, None)

    def matrix_from_csr_lower(self, n, rowptr, colind, vals):
        """half-stored (sym=1) host matrix -> full device matrix"""
        rowptr, colind, vals = _i32(rowptr), _i32(colind), _f64(vals)
        out = C.c_void_p()
        _ck(lib().ffcuda_matrix_from_csr_lower(_h(self), int(n), C.c_int64(len(colind)), _p(rowptr), _p(colind), _p(vals), C.byref(out)),
            self.h)
        return Matrix(out.value, self, None)

    def matrix_from_csr(self, n, rowptr, colind, vals):
        rowptr, colind, vals = _i32(rowptr), _i32(colind), _f64(vals)
        out = C.c_void_p()
        _ck(lib().ffcuda_matrix_from_csr(_h(self), int(n), C.c_int64(len(colind)), _p(rowptr), _p(colind), _p(vals), C.byref(out)),
            self.h)
        return Matrix(out.value, self, None)


class Mesh(_Handle):
    _destroy = "ffcuda_mesh_destroy"

    def info(self):
        d, nv, nt, nbe = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _ck(lib().ffcuda_mesh_info(_h(self), C.byref(d), C.byref(nv), C.byref(nt), C.byref(nbe)), self.ctx.h)
        return d.value, nv.value, nt.value, nbe.value

    def download(self):
        dim, nv, nt, nbe = self.info()
        m = dict(dim=dim, xyz=np.zeros((nv, dim)), conn=np.zeros((nt, dim + 1), np.int32), elab=np.zeros(nt, np.int32),
                 bconn=np.zeros((nbe, dim), np.int32), blab=np.zeros(nbe, np.int32), belem=np.zeros(nbe, np.int32),
                 bface=np.zeros(nbe, np.int32))
        _ck(lib().ffcuda_mesh_download(_h(self), _p(m["xyz"]), _p(m["conn"]), _p(m["elab"]), _p(m["bconn"]), _p(m["blab"]),
                                       _p(m["belem"]), _p(m["bface"])), self.ctx.h)
        return m

    def fe_table(self, order, dofs, table, qpts, op=0, border=False, e2n=None, dstride=1, doff=0, scale=1.0, labels=None, offset=0,
                 accumulate=False):
        """table[offset + u*nq + q] (+)= scale * d^op f(P_q of unit u) for the P0/P1/P2 function whose dofs are the device Vec `dofs`"""
        qpts, e2n, lab = _f64(qpts), _i32(e2n), _i32(labels)
        dim = self.info()[0]
        nq = qpts.size // (dim - 1 if border else dim)
        _ck(lib().ffcuda_fe_table(_h(self), int(order), _p(e2n), int(dstride), int(doff), _h(dofs), int(op), int(bool(border)), int(nq),
                                  _p(qpts), C.c_double(scale), 0 if lab is None else len(lab), _p(lab), _h(table), C.c_int64(offset),
                                  int(accumulate)), self.ctx.h)

    def adjacency(self):
        """element adjacency as GenericMesh::BuildAdj numbers it: adj[(dim+1)*k+i] = (dim+1)*k'+i' | -1 (boundary) | -2"""
        dim, _, nt, _ = self.info()
        adj = np.zeros((dim + 1) * nt, np.int32)
        _ck(lib().ffcuda_mesh_adjacency(_h(self), _p(adj), None), self.ctx.h)
        return adj

    def buildlayers(self, nlayer, ni=None, zmin=None, zmax=None, regmap=(), midmap=(), upmap=(), downmap=()):
        """`buildlayers(Th2, nlayer, ...)` on the device (self = the 2-D mesh); maps are flat (old,new,...) sequences"""
        ni, zmin, zmax = _i32(ni), _f64(zmin), _f64(zmax)
        maps = [np.ascontiguousarray(np.asarray(m_, dtype=np.int32).ravel()) for m_ in (regmap, midmap, upmap, downmap)]
        margs = []
        for m_ in maps:
            margs += [int(m_.size // 2), _p(m_) if m_.size else None]
        out = C.c_void_p()
        _ck(lib().ffcuda_mesh_buildlayers(_h(self), int(nlayer), _p(ni), _p(zmin), _p(zmax), *margs, C.byref(out)), self.ctx.h)
        return Mesh(out.value, self.ctx)

    def local_to_global(self):
        no, nl = C.c_int(), C.c_int()
        _ck(lib().ffcuda_mesh_local_to_global(_h(self), C.byref(no), C.byref(nl), None), self.ctx.h)
        gid = np.zeros(nl.value, np.int64)
        _ck(lib().ffcuda_mesh_local_to_global(_h(self), C.byref(no), C.byref(nl), _p(gid)), self.ctx.h)
        return no.value, gid

    def space_distributed(self, order, ncomp, elem2node, nowned, nlocal, nbr, recv_off, recv_cnt, send_ptr, send_idx):
        """a space with its own node table and node-level halo lists on a distributed mesh (P2): ffcuda_space_create_distributed"""
        e2n, nbr, ro, rc, sp, si = (_i32(a) for a in (elem2node, nbr, recv_off, recv_cnt, send_ptr, send_idx))
        out = C.c_void_p()
        _ck(lib().ffcuda_space_create_distributed(_h(self), order, ncomp, _p(e2n), int(nowned), int(nlocal), len(nbr), _p(nbr), _p(ro), _p(rc),
                                                  _p(sp), _p(si), C.byref(out)), self.ctx.h)
        return Space(out.value, self.ctx, self)

    def space(self, order=1, ncomp=1, elem2node=None, nnodes=0):
        e2n = _i32(elem2node)
        out = C.c_void_p()
        _ck(lib().ffcuda_space_create(_h(self), order, ncomp, _p(e2n), int(nnodes), C.byref(out)), self.ctx.h)
        return Space(out.value, self.ctx, self)


class Space(_Handle):
    _destroy = "ffcuda_space_destroy"

    def __init__(self, h, ctx, mesh):
        super().__init__(h, ctx)
        self.mesh = mesh

    def info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        _ck(lib().ffcuda_space_info(_h(self), C.byref(a), C.byref(b), C.byref(c)), self.ctx.h)
        return a.value, b.value, c.value  # ndof, ndofK, nnodes

    def dofs(self):
        ndof, ndofK, _ = self.info()
        nt = self.mesh.info()[2]
        d = np.zeros((nt, ndofK), np.int32)
        _ck(lib().ffcuda_space_download_dofs(_h(self), _p(d)), self.ctx.h)
        return d

    def symbolic(self):
        out = C.c_void_p()
        _ck(lib().ffcuda_symbolic(_h(self), C.byref(out)), self.ctx.h)
        return Pattern(out.value, self.ctx, self)

    def assemble_rect(self, unknown_space, terms, qpts, qw, labels=None):
        """`matrix B = vb(Uh,Vh)`: self = the test space Vh (rows), unknown_space = Uh (columns), both on the same device mesh"""
        arr = (BTerm * max(len(terms), 1))()
        for k, (uc, uo, vc, vo, c) in enumerate(terms):
            arr[k] = BTerm(uc, uo, vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        out = C.c_void_p()
        _ck(lib().ffcuda_assemble_bilinear_rect(_h(self), _h(unknown_space), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                                0 if lab is None else len(lab), _p(lab), C.byref(out)), self.ctx.h)
        return Matrix(out.value, self.ctx, None)

    def bc_from_labels(self, labels, compmask, values):
        labels, values = _i32(labels), _f64(values)
        out = C.c_void_p()
        _ck(lib().ffcuda_bc_from_labels(_h(self), len(labels), _p(labels), int(compmask), _p(values), C.byref(out)), self.ctx.h)
        return BC(out.value, self.ctx)

    def bc_from_pairs(self, dofs, vals):
        dofs, vals = _i32(dofs), _f64(vals)
        out = C.c_void_p()
        _ck(lib().ffcuda_bc_from_pairs(_h(self), len(dofs), _p(dofs), _p(vals), C.byref(out)), self.ctx.h)
        return BC(out.value, self.ctx)

    def assemble_linear(self, b, terms, qpts, qw, labels=None, accumulate=False):
        arr = (LTerm * max(len(terms), 1))()
        for k, (vc, vo, c) in enumerate(terms):
            arr[k] = LTerm(vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        _ck(lib().ffcuda_assemble_linear(_h(b), _h(self), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                         0 if lab is None else len(lab), _p(lab), int(accumulate)), self.ctx.h)

    def assemble_linear_qvalues(self, b, qpts, qw, fq, accumulate=False):
        """b (+)= int(f v) with f given at the quadrature nodes: fq[c, k, q] (ncomp x nt x nq)"""
        qpts, qw = _f64(qpts), _f64(qw)
        fq, pfq = _table(fq)
        _ck(lib().ffcuda_assemble_linear_qvalues(_h(b), _h(self), len(qw), _p(qpts), _p(qw), pfq, int(accumulate)), self.ctx.h)

    def assemble_linear_qterms(self, b, qpts, qw, fq, accumulate=False):
        """b (+)= int(sum_s f_s d^s v) with the f_s given at the quadrature nodes: fq[c, s, k, q] (ncomp x (dim+1) x nt x nq)"""
        qpts, qw = _f64(qpts), _f64(qw)
        fq, pfq = _table(fq)
        _ck(lib().ffcuda_assemble_linear_qterms(_h(b), _h(self), len(qw), _p(qpts), _p(qw), pfq, int(accumulate)), self.ctx.h)

    def assemble_linear_boundary_qvalues(self, b, qpts, qw, gq, accumulate=True):
        """b (+)= boundary integral of g v with g given at the face quadrature nodes: gq[c, ib, q] (0 where it does not go)"""
        qpts, qw = _f64(qpts), _f64(qw)
        gq, pgq = _table(gq)
        _ck(lib().ffcuda_assemble_linear_boundary_qvalues(_h(b), _h(self), len(qw), _p(qpts), _p(qw), pgq, int(accumulate)), self.ctx.h)

    def assemble_linear_boundary(self, b, terms, qpts, qw, labels=None, accumulate=True):
        """b (+)= int2d(Th3, labels)(c v) / int1d(Th, labels)(c v); qpts: nq x (dim-1) face reference coordinates"""
        arr = (LTerm * max(len(terms), 1))()
        for k, (vc, vo, c) in enumerate(terms):
            arr[k] = LTerm(vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        _ck(lib().ffcuda_assemble_linear_boundary(_h(b), _h(self), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                                  0 if lab is None else len(lab), _p(lab), int(accumulate)), self.ctx.h)


class Pattern(_Handle):
    _destroy = "ffcuda_pattern_destroy"

    def __init__(self, h, ctx, space):
        super().__init__(h, ctx)
        self.space = space

    def info(self):
        n, nnz = C.c_int(), C.c_int64()
        _ck(lib().ffcuda_pattern_info(_h(self), C.byref(n), C.byref(nnz)), self.ctx.h)
        return n.value, nnz.value

    def download(self, rp=None, ci=None):
        n, nnz = self.info()
        rp = np.zeros(n + 1, np.int32) if rp is None else rp
        ci = np.zeros(nnz, np.int32) if ci is None else ci
        assert rp.dtype == np.int32 and ci.dtype == np.int32 and len(rp) == n + 1 and len(ci) == nnz
        _ck(lib().ffcuda_pattern_download(_h(self), _p(rp), _p(ci)), self.ctx.h)
        return rp, ci

    def download_lower(self):
        """(rowptr, colind) of the lower triangle: what a sym=1 MatriceMorse holds"""
        n, _ = self.info()
        nl = C.c_int64()
        _ck(lib().ffcuda_pattern_lower_nnz(_h(self), C.byref(nl)), self.ctx.h)
        rp, ci = np.zeros(n + 1, np.int32), np.zeros(nl.value, np.int32)
        _ck(lib().ffcuda_pattern_download_lower(_h(self), _p(rp), _p(ci)), self.ctx.h)
        return rp, ci

    def download_async(self, rp, ci):
        """copies behind the work enqueued so far, on a second stream; rp / ci (pinned numpy arrays) are valid after ctx.sync()"""
        _ck(lib().ffcuda_pattern_download_async(_h(self), _p(rp), _p(ci)), self.ctx.h)

    def matrix(self):
        out = C.c_void_p()
        _ck(lib().ffcuda_matrix_create(_h(self), C.byref(out)), self.ctx.h)
        return Matrix(out.value, self.ctx, self)


class Matrix(_Handle):
    _destroy = "ffcuda_matrix_destroy"

    def __init__(self, h, ctx, pattern):
        super().__init__(h, ctx)
        self.pattern = pattern

    def info(self):
        n, nnz = C.c_int(), C.c_int64()
        _ck(lib().ffcuda_matrix_info(_h(self), C.byref(n), C.byref(nnz)), self.ctx.h)
        return n.value, nnz.value

    def download(self, out=None):
        _, nnz = self.info()
        v = np.zeros(nnz) if out is None else out
        _ck(lib().ffcuda_matrix_download(_h(self), _p(v)), self.ctx.h)
        return v

    def shape(self):
        n, m, nnz = C.c_int(), C.c_int(), C.c_int64()
        _ck(lib().ffcuda_matrix_shape(_h(self), C.byref(n), C.byref(m), C.byref(nnz)), self.ctx.h)
        return n.value, m.value, nnz.value

    def download_csr(self):
        """(rowptr, colind, vals) of a matrix without a pattern object"""
        n, _, nnz = self.shape()
        rp, ci, v = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        _ck(lib().ffcuda_matrix_download_csr(_h(self), _p(rp), _p(ci), _p(v)), self.ctx.h)
        return rp, ci, v

    def download_lower(self):
        nl = C.c_int64()
        _ck(lib().ffcuda_pattern_lower_nnz(_h(self.pattern), C.byref(nl)), self.ctx.h)
        out = np.zeros(nl.value)
        _ck(lib().ffcuda_matrix_download_lower(_h(self), _p(out)), self.ctx.h)
        return out

    def cg_stop_threshold(self):
        v = C.c_double()
        _ck(lib().ffcuda_cg_stop_threshold(_h(self), C.byref(v)), self.ctx.h)
        return v.value

    def export_device(self):
        """borrowed device pointers (ints) to the CSR triple: (rowptr, colind, vals, n, nnz)"""
        rp, ci, va = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n, nnz = C.c_int(), C.c_int64()
        _ck(lib().ffcuda_matrix_export_device(_h(self), C.byref(rp), C.byref(ci), C.byref(va), C.byref(n), C.byref(nnz)), self.ctx.h)
        return rp.value, ci.value, va.value, n.value, nnz.value

    def download_coo(self, index_base=0):
        """the triple of `[I,J,C] = A`: row indices expanded on the device"""
        _, nnz = self.info()
        I, J, V = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        _ck(lib().ffcuda_matrix_download_coo(_h(self), _p(I), _p(J), _p(V), int(index_base)), self.ctx.h)
        return I, J, V

    def write_morse(self, path, half=False):
        """FreeFEM's Morse text format (`ofstream << A` after A.CSR) straight from the device CSR"""
        _ck(lib().ffcuda_matrix_write_morse(_h(self), os.fsencode(path), 1 if half else 0), self.ctx.h)

    def upload(self, vals):
        vals = _f64(vals)
        _ck(lib().ffcuda_matrix_upload(_h(self), _p(vals)), self.ctx.h)

    def assemble(self, terms, qpts, qw, labels=None, accumulate=False):
        arr = (BTerm * max(len(terms), 1))()
        for k, (uc, uo, vc, vo, c) in enumerate(terms):
            arr[k] = BTerm(uc, uo, vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        _ck(lib().ffcuda_assemble_bilinear(_h(self), _h(self.pattern.space), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                           0 if lab is None else len(lab), _p(lab), int(accumulate)), self.ctx.h)

    def assemble_qcoef(self, terms, qpts, qw, cq, accumulate=False):
        """A (+)= the form multiplied by the coefficient whose values at the quadrature nodes are cq[k, q] (P1 spaces)"""
        arr = (BTerm * max(len(terms), 1))()
        for k, (uc, uo, vc, vo, c) in enumerate(terms):
            arr[k] = BTerm(uc, uo, vc, vo, c)
        qpts, qw = _f64(qpts), _f64(qw)
        cq, pcq = _table(cq)
        _ck(lib().ffcuda_assemble_bilinear_qcoef(_h(self), _h(self.pattern.space), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                                 pcq, int(accumulate)), self.ctx.h)

    def assemble_boundary_qcoef(self, terms, qpts, qw, cq, labels=None, accumulate=True):
        """A (+)= boundary integral of alpha u v with alpha given at the face quadrature nodes: cq[ib, q]"""
        arr = (BTerm * max(len(terms), 1))()
        for k, (uc, uo, vc, vo, c) in enumerate(terms):
            arr[k] = BTerm(uc, uo, vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        cq, pcq = _table(cq)
        _ck(lib().ffcuda_assemble_bilinear_boundary_qcoef(_h(self), _h(self.pattern.space), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                                          pcq, 0 if lab is None else len(lab), _p(lab), int(accumulate)), self.ctx.h)

    def assemble_boundary(self, terms, qpts, qw, labels=None, accumulate=True):
        """A (+)= int2d(Th3, labels)(c u v) / int1d(Th, labels)(c u v) (Robin terms); qpts: nq x (dim-1) face coordinates"""
        arr = (BTerm * max(len(terms), 1))()
        for k, (uc, uo, vc, vo, c) in enumerate(terms):
            arr[k] = BTerm(uc, uo, vc, vo, c)
        qpts, qw, lab = _f64(qpts), _f64(qw), _i32(labels)
        _ck(lib().ffcuda_assemble_bilinear_boundary(_h(self), _h(self.pattern.space), len(terms), arr, len(qw), _p(qpts), _p(qw),
                                                    0 if lab is None else len(lab), _p(lab), int(accumulate)), self.ctx.h)

    def apply_bc(self, bc, tgv=1e30):
        _ck(lib().ffcuda_matrix_apply_bc(_h(self), _h(bc), C.c_double(tgv)), self.ctx.h)

    def spmv(self, x, y):
        _ck(lib().ffcuda_spmv(_h(self), _h(x), _h(y)), self.ctx.h)

    def cg(self, b, x, eps=1e-6, itmax=0, tgv=1e30):
        it, conv, g = C.c_int(), C.c_int(), C.c_double()
        _ck(lib().ffcuda_cg(_h(self), _h(b), _h(x), C.c_double(eps), int(itmax), C.c_double(tgv), C.byref(it), C.byref(conv),
                            C.byref(g)), self.ctx.h)
        return it.value, conv.value, g.value

    def cg_host(self, b, x, eps=1e-6, itmax=0, tgv=1e30):
        """b, x: contiguous float64 numpy arrays (x: initial guess in, solution out)."""
        assert b.dtype == np.float64 and x.dtype == np.float64 and b.flags.c_contiguous and x.flags.c_contiguous
        it, conv, g = C.c_int(), C.c_int(), C.c_double()
        _ck(lib().ffcuda_cg_host(_h(self), _p(b), _p(x), C.c_double(eps), int(itmax), C.c_double(tgv), C.byref(it),
                                 C.byref(conv), C.byref(g)), self.ctx.h)
        return it.value, conv.value, g.value

    def gmres(self, b, x, eps=1e-6, itmax=0, restart=0, tgv=1e30):
        """SolverGMRES of FreeFEM on the device (right Jacobi, modified Gram-Schmidt); returns (iterations, converged, relres)"""
        it, conv, r = C.c_int(), C.c_int(), C.c_double()
        _ck(lib().ffcuda_gmres(_h(self), _h(b), _h(x), C.c_double(eps), int(itmax), int(restart), C.c_double(tgv), C.byref(it),
                               C.byref(conv), C.byref(r)), self.ctx.h)
        return it.value, conv.value, r.value

    def gmres_host(self, b, x, eps=1e-6, itmax=0, restart=0, tgv=1e30):
        assert b.dtype == np.float64 and x.dtype == np.float64 and b.flags.c_contiguous and x.flags.c_contiguous
        it, conv, r = C.c_int(), C.c_int(), C.c_double()
        _ck(lib().ffcuda_gmres_host(_h(self), _p(b), _p(x), C.c_double(eps), int(itmax), int(restart), C.c_double(tgv),
                                    C.byref(it), C.byref(conv), C.byref(r)), self.ctx.h)
        return it.value, conv.value, r.value


class Vec(_Handle):
    _destroy = "ffcuda_vec_destroy"

    def __init__(self, h, ctx, n):
        super().__init__(h, ctx)
        self.n = n

    def upload(self, host):
        host = _f64(host)
        assert len(host) == self.n
        _ck(lib().ffcuda_vec_upload(_h(self), _p(host)), self.ctx.h)

    def download(self, out=None):
        v = np.zeros(self.n) if out is None else out
        _ck(lib().ffcuda_vec_download(_h(self), _p(v)), self.ctx.h)
        return v

    def fill(self, value):
        _ck(lib().ffcuda_vec_fill(_h(self), C.c_double(value)), self.ctx.h)

    def ptr(self):
        return lib().ffcuda_vec_ptr(_h(self))

    def apply_bc(self, bc, tgv=1e30):
        _ck(lib().ffcuda_vec_apply_bc(_h(self), _h(bc), C.c_double(tgv)), self.ctx.h)

    def set_bc_values(self, bc):
        _ck(lib().ffcuda_vec_set_bc_values(_h(self), _h(bc)), self.ctx.h)


class BC(_Handle):
    _destroy = "ffcuda_bc_destroy"

    def count(self):
        n = C.c_int()
        _ck(lib().ffcuda_bc_count(_h(self), C.byref(n)), self.ctx.h)
        return n.value
