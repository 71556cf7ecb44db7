// b200_opgen_lean.cpp -- the "lean" fused kernel for gradient-free (mass-like) operators: BP1 / BP2 and every user operator of
// the same shape (one active input group and one active output group on the same tensor basis with EVAL_INTERP only, any number
// of strided EVAL_NONE inputs and EVAL_WEIGHT, an arbitrary user QFunction).  QFunction layout id 4 of the kernel shape.
//
// Same contract as the general generator (b200_opgen.cpp; it replaces backends/cuda-gen/ceed-cuda-gen-operator-build.cpp:1158-1681
// for this operator class), different cost structure.  At low order the general kernel is bound by instruction issue and
// shared-memory wavefronts, not by HBM (ncu, profiles/r02_ncu_full_summary.txt: BP1 p=3 = 382 warp instructions and 100
// shared-memory wavefronts per 1.7 KB element, 54 % issue-active).  What this kernel does about it:
//
//   * ONE IN-PLACE PLANE per (element, component): Q x Q x P doubles laid out [qz][qy | j][i].  The z-contraction writes rows
//     j < P, the y-contraction expands every (i, qz) line IN PLACE (the lane that owns the line reads its P values and writes its
//     Q values), the x-line stage reads the P values of an x-line, runs X, the QFunction on the Q points of the line and X^T in
//     registers, and writes P values back; Y^T and Z^T walk the same plane backwards.  800 B per element at p = 3 instead of 2000.
//   * A warp owns E (typically 8) elements per iteration and nothing is shared between warps (__syncwarp() only): every stage is
//     straight-line code over ceil(tasks / 32) fully unrolled rounds, so the per-stage overhead (calls, address set-up, constant
//     loads) is paid once per E elements and the lanes are nearly full (16 E, 20 E and 25 E tasks at p = 3).
//   * All global-memory index streams of a batch are read up front: the gather stage loads the element offsets AND the scatter
//     targets of all its rounds before the first dependent load and parks the targets in a per-lane shared-memory slot, so the
//     scatter stage at the end of the batch never waits on a table load.  Quadrature data is loaded one round ahead of its use.
//   * x-lines of an even P are moved with 16-byte shared-memory accesses.
//   * Stage bit 32 = BULK PIPELINE: the three contiguous streams of a batch -- element offsets, scatter targets and the quadrature
//     data of the E elements -- are fetched by the copy engine (cp.async.bulk + mbarrier, one copy per stream and batch issued by
//     lane 0) into single shared-memory buffers, each re-armed for the warp's NEXT batch the moment its consumer stage is done:
//     offsets after the gather stage, quadrature data after the x-line stage, targets after the scatter stage.  Every stream is
//     then in flight for most of a batch period with no registers held and no load instructions issued; the only demand loads
//     left are the gathers of u.  Sources are aligned down to 16 bytes at run time (readers add the shift).
//   * Stage bit 128 = IN-KERNEL FINALIZE: the elements are cut into K contiguous parts (batch-aligned); a warp that has finished its
//     batches of part c publishes that (release add on a per-part counter), and -- once every warp has done so for part c - 1 --
//     folds its static slice of the shared nodes part c - 1 completed into v (the work of k_halo_finalize), while the halo values and
//     the owner values it reads are still in L2.  One launch instead of two, no second pass over DRAM for the halo buffer.  All CTAs
//     must be resident (cooperative launch); a warp only ever waits for parts that lie one part duration in the past.
//   * CEED_B200_RUNS (experimental, off): run scatter -- every warp owns a contiguous run of elements and adds E-entries whose earlier
//     touchers it processed itself straight into v (B200RunScatter).  Bitwise equal results, finalize pass 20-30 % shorter, but the
//     fused kernel loses its streaming locality (2x slower on B200): kept as a tested option, see DESIGN.md.
//   * Stage bit 1 = ELEMENT-INTERLEAVED COLUMNS: the gather + Z and Z^T + scatter stages enumerate their node columns as (i, element, j)
//     instead of (i, j, element).  The E elements of a batch are x-neighbours on a lexicographically ordered mesh, so the lanes of one
//     load / store then walk ONE node row across the whole batch (E p + 1 consecutive L-nodes, 200 B at p = 3, E = 8) instead of eight
//     32-byte row segments in eight different 128-byte lines: every request of the gather and of the owner / halo stores touches 2-3 lines
//     instead of 8-10.  The L1 data pipe processes one line per wavefront, and it is the unit that bounds this kernel (ncu: 91 % busy).
//     The element-major offset and target tables would pay for that (a request would touch every element's block): the warp first copies
//     the batch's two tables into shared memory with coalesced vector loads (element pitch = P (mod 32) ints: the interleaved reads are
//     conflict-free) and both column stages read them from there -- which also replaces the parked targets (scripts/model/lean_columns.py).
//     Stage bit 2 pads the element stride of the planes to P (mod 16) doubles, which keeps those stages free of bank conflicts for every P.
//   * Stage bit 4 = 16-BYTE QUADRATURE-DATA LOADS: a lane of the x-line stage reads Q consecutive doubles per component; as 8-byte loads
//     with a lane stride of 8 Q bytes every one of the Q requests touches the same ~Q / 4 lines per 8 lanes again.  With this bit the
//     lane reads the 16-byte aligned window around its line with ceil(Q / 2) 16-byte loads (the alignment is taken from the address at
//     run time; the first / last line of an array and misaligned user arrays take scalar loads) and, for an odd Q, selects.
//   * Full batches run code without any tail clamps; the (at most one) partial batch of a launch runs a second instantiation.
#include <algorithm>
#include <cstdlib>
#include <sstream>

#include "b200_opgen.h"

using std::string;
typedef std::ostringstream oss;

namespace {

string hexd(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%a", v);
  return buf;
}
string S(long long v) { return std::to_string(v); }

struct LeanGen {
  B200Operator op;
  B200OpPlan  *plan;
  int          add;
  oss          c;
  int          P, Q, E, nc, PS, ES, P2, P3, bid;
  bool         staged = false;  // bulk pipeline: st_idx (stage bit 8: offsets + targets) and / or st_qd (bit 32: quadrature data)
  bool         st_idx = false, st_qd = false;
  bool         fin   = false;   // stage bit 128: in-kernel finalize of the deterministic scatter (whole-mesh launches)
  bool         pf_qd = false;   // stage bit 64: bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the directly loaded quadrature data of the NEXT batch
  bool         vqd    = false;  // stage bit 4: directly loaded quadrature data of an x-line through 16-byte loads (aligned window)
  bool         ilv    = false;  // stage bit 1: element-interleaved node columns in the gather / scatter stages
  bool         runs   = false;  // run scatter (experimental, CEED_B200_RUNS): per-warp element runs, read-modify-write entries
  const B200GenGroup *gin, *gout;
  bool qd_staged(const B200GenField &fd) const { return st_qd && fd.emode == B200_EVAL_NONE && fd.qd_off >= 0; }
  // stage bit 4: the Q consecutive doubles a lane reads per (x-line, component) come in as ceil(Q / 2) 16-byte loads from the 16-byte
  // aligned window around them (an odd Q makes every other line start 8 bytes off: those lanes pick their values one double later)
  bool qd_vector(const B200GenField &fd) const { return vqd && !qd_staged(fd) && fd.emode == B200_EVAL_NONE && fd.rstr->is_strided && fd.rstr->strides[0] == 1 && Q >= 2; }
  int  qd_nv() const { return (Q + 1) / 2; }
  // value q of component cc of EVAL_NONE input f in round x
  string qd_value(size_t f, int cc, int q, const string &x) const {
    const B200GenField &fd = plan->in_fields[f];
    if (!qd_vector(fd)) return "d" + S((long long)f) + "_" + S(cc * Q + q) + x;
    const string w = "w" + S((long long)f) + "_" + S(cc) + "_";
    auto at = [&](int j) { return w + S(j / 2) + x + (j % 2 ? ".y" : ".x"); };
    if (Q % 2 == 0) return at(q);
    return "(od" + S((long long)f) + "_" + S(cc) + x + " ? " + at(q + 1) + " : " + at(q) + ")";
  }

  void contract(const string &mat, int n_in, int n_out, bool transposed, const string &in, const string &out, const string &ind) {
    for (int o = 0; o < n_out; o++) {
      for (int i = 0; i < n_in; i++) {
        const int idx = transposed ? i * n_out + o : o * n_in + i;
        if (i == 0) c << ind << "double " << out << o << " = " << mat << "[" << idx << "] * " << in << i << ";\n";
        else c << ind << out << o << " = fma(" << mat << "[" << idx << "], " << in << i << ", " << out << o << ");\n";
      }
    }
  }

  void emit_header() {
    B200QFunction qf = op->qf;
    c << "// Fused operator kernel (lean, in-place plane) generated by ceed-b200 for QFunction " << qf->kernel_name << "\n";
    c << "#define CEED_Q_VLA " << Q << "  // the QFunction is called on whole x-lines (Q = " << Q << " points)\n";
    c << "#include <b200-jit.h>\n";
    if (staged || pf_qd) c << "#include <b200-tma.h>\n";
    c << "#include \"" << qf->source_path << "\"\n\n";
    for (size_t b = 0; b < plan->bases.size(); b++) {
      const B200Basis bs = plan->bases[b].basis;
      auto emit_mat      = [&](const string &name, const std::vector<double> &m) {
        c << "__constant__ double " << name << "[" << m.size() << "] = {";
        for (size_t i = 0; i < m.size(); i++) c << (i ? ", " : "") << hexd(m[i]);
        c << "};\n";
      };
      emit_mat("cB" + S(b), bs->interp);
      emit_mat("cW" + S(b), bs->q_weight);
    }
    c << "\nstruct B200OpArgs {\n  long long num_elem;\n  void *ctx;\n  const double *in_ptr[16];\n  double *out_ptr[16];\n"
      << "  const int *in_idx[16];\n  const int *out_idx[16];\n  double *out_aux[16];\n"
      << "  const int *ord_pred_ptr, *ord_pred_idx;\n  int *ord_flags, *ord_sync;\n  long long ord_num_halo;\n};\n"
      << "__constant__ B200OpArgs b200a;\n\n";
    c << "extern __shared__ __align__(16) double sm[];\n";
    c << "__shared__ long long b200_ne;\n\n";
    c << "#define B200_SMW (sm + (threadIdx.x >> 5) * " << plan->group_smem_bytes / 8 << ")\n";
    c << "#define B200_TGS ((int *)(B200_SMW + " << plan->lean_tg_off / 8 << "))\n";
    if (ilv) c << "#define B200_OFS ((int *)(B200_SMW + " << plan->lean_off_off / 8 << "))\n";
    if (staged) {
      if (st_idx) c << "#define B200_OFS ((int *)(B200_SMW + " << plan->lean_off_off / 8 << "))\n";
      c << "#define B200_BAR(s) (b200_smem_u32((char *)B200_SMW + " << plan->mbar_off << " + 8 * (s)))  // 0 offsets, 1 quadrature data, 2 targets\n";
      // run-time shift (in elements) of a bulk-copied block whose source was aligned down to 16 bytes
      c << "#define B200_SHIFT(ptr, esize) ((int)(((unsigned long long)(ptr) & 15ULL) / (esize)))\n";
      // Bytes of the bulk copy of [src, src + bytes): the source is aligned down to 16 bytes and the size rounded up to 16 -- except in
      // the last block of the array (`end` = one past its last byte), where rounding up would read past the allocation: the copy then
      // stops at the last 16-byte boundary inside the array and the few trailing elements are moved by the issuing lane itself.
      c << "static __device__ __forceinline__ unsigned b200_lean_nb(const void *src, unsigned bytes, const void *end) {\n"
        << "  const unsigned long long a0 = (unsigned long long)src & ~15ULL;\n  unsigned nb = (unsigned)(((unsigned long long)src & 15ULL) + bytes + 15u) & ~15u;\n"
        << "  if (a0 + nb > (unsigned long long)end) nb = (unsigned)((unsigned long long)end - a0) & ~15u;\n  return nb;\n}\n";
      c << "template <typename T> static __device__ __forceinline__ void b200_lean_copy(void *dst, const T *src, unsigned bytes, unsigned bar, const T *end) {\n"
        << "  const unsigned long long a0 = (unsigned long long)src & ~15ULL;\n  const unsigned nb = b200_lean_nb(src, bytes, end);\n"
        << "  for (const T *p = (const T *)(a0 + nb) < src ? src : (const T *)(a0 + nb); p < (const T *)((const char *)src + bytes); p++)\n"
        << "    *(T *)((char *)dst + ((unsigned long long)p - a0)) = *p;  // (only ever runs for the last block of the array)\n"
        << "  if (nb) b200_bulk_g2s(b200_smem_u32(dst), (const void *)a0, nb, bar);\n}\n";
      // issue functions (lane 0 of the warp; ne = elements of the batch starting at e0)
      if (st_idx)
      c << "static __device__ __noinline__ void b200_lean_issue_off(const long long e0, const int ne) {\n"
        << "  const int *src = b200a.in_idx[" << gin->slot << "] + e0 * " << P3 << ";\n  const unsigned bar = B200_BAR(0);\n"
        << "  const int *end = b200a.in_idx[" << gin->slot << "] + b200a.num_elem * " << P3 << ";\n"
        << "  b200_fence_proxy_async();\n  b200_mbar_expect_tx(bar, b200_lean_nb(src, ne * " << P3 * 4 << ", end));\n"
        << "  b200_lean_copy(B200_OFS, src, ne * " << P3 * 4 << ", bar, end);\n}\n";
      if (st_idx)
      c << "static __device__ __noinline__ void b200_lean_issue_tg(const long long e0, const int ne) {\n"
        << "  const int *src = b200a.out_idx[" << gout->slot << "] + e0 * " << P3 << ";\n  const unsigned bar = B200_BAR(2);\n"
        << "  const int *end = b200a.out_idx[" << gout->slot << "] + b200a.num_elem * " << P3 << ";\n"
        << "  b200_fence_proxy_async();\n  b200_mbar_expect_tx(bar, b200_lean_nb(src, ne * " << P3 * 4 << ", end));\n"
        << "  b200_lean_copy(B200_TGS, src, ne * " << P3 * 4 << ", bar, end);\n}\n";
      bool any_qd = false;
      for (auto &fd : plan->in_fields) any_qd = any_qd || qd_staged(fd);
      if (any_qd) {
        const int Q3 = Q * Q * Q;
        c << "static __device__ __noinline__ void b200_lean_issue_qd(const long long e0, const int ne) {\n  const unsigned bar = B200_BAR(1);\n";
        int k = 0;
        for (auto &fd : plan->in_fields) {
          if (!qd_staged(fd)) continue;
          for (int cc = 0; cc < fd.nc; cc++, k++) {
            c << "  const double *src" << k << " = b200a.in_ptr[" << fd.slot << "] + " << (long long)cc * fd.rstr->strides[1] << "LL + e0 * " << Q3 << "LL;\n";
            c << "  const double *end" << k << " = b200a.in_ptr[" << fd.slot << "] + " << (long long)cc * fd.rstr->strides[1] << "LL + b200a.num_elem * " << Q3 << "LL;\n";
          }
        }
        c << "  b200_fence_proxy_async();\n  b200_mbar_expect_tx(bar, 0u";
        for (int i = 0; i < k; i++) c << " + b200_lean_nb(src" << i << ", ne * " << Q3 * 8 << ", end" << i << ")";
        c << ");\n";
        k = 0;
        for (auto &fd : plan->in_fields) {
          if (!qd_staged(fd)) continue;
          for (int cc = 0; cc < fd.nc; cc++, k++)
            c << "  b200_lean_copy((char *)B200_SMW + " << fd.qd_off + cc * fd.qd_cs * 8 << ", src" << k << ", ne * " << Q3 * 8 << ", bar, end" << k << ");\n";
        }
        c << "}\n";
      }
    }
    if (pf_qd) {
      const int Q3 = Q * Q * Q;
      c << "static __device__ __noinline__ void b200_lean_prefetch_qd(const long long e0, const int ne) {\n";
      for (auto &fd : plan->in_fields) {
        const bool contiguous = fd.emode == B200_EVAL_NONE && fd.rstr->is_strided && fd.rstr->strides[0] == 1 && fd.rstr->strides[2] == fd.rstr->elem_size;
        if (!contiguous) continue;
        for (int cc = 0; cc < fd.nc; cc++) {
          c << "  { const unsigned long long a = (unsigned long long)(b200a.in_ptr[" << fd.slot << "] + " << (long long)cc * fd.rstr->strides[1] << "LL + e0 * " << Q3 << "LL);\n";
          c << "    const unsigned long long end = (unsigned long long)(b200a.in_ptr[" << fd.slot << "] + " << (long long)cc * fd.rstr->strides[1] << "LL + b200a.num_elem * " << Q3
            << "LL);\n";
          c << "    unsigned nb = (unsigned)(((a & 15ULL) + (unsigned long long)ne * " << Q3 * 8 << " + 15) & ~15ULL);\n";
          c << "    if ((a & ~15ULL) + nb > end) nb = (unsigned)(end - (a & ~15ULL)) & ~15u;  // never past the end of the array\n";
          c << "    if (nb) b200_bulk_prefetch_l2((const void *)(a & ~15ULL), nb); }\n";
        }
      }
      c << "}\n";
    }
    c << "\n";
  }

  // element number of local element `le` (tail batches clamp to the last element of the launch)
  string elem(const string &le) const { return "(TAIL ? ((e0 + " + le + " * st < lim) ? e0 + " + le + " * st : e0) : e0 + " + le + " * st)"; }
  // trailing parameters of the stage functions: element stride inside a batch and end of the warp's element range (run mode), or the
  // parity of the mbarrier phase (bulk pipeline: batches are contiguous)
  string stage_params() const { return staged ? ", const int par" : (runs ? ", const int st, const long long lim" : ""); }
  string stage_locals() const { return runs ? "" : "  const int st = 1;\n  const long long lim = b200_ne;\n  (void)st; (void)lim;\n"; }
  string stage_args() const { return staged ? ", par" : (runs ? ", st, lim" : ""); }

  // vector width (ints) of the table staging loads of stage bit 1: element blocks and the padded pitch must be multiples of it
  int ilv_width() const {
    for (int w : {4, 2})
      if (P3 % w == 0 && plan->lean_ip % w == 0) return w;
    return 1;
  }
  // node column (ij = i + P j, local element le) of task `t` of the gather / scatter stages: (i, j, element) by default, (i, element, j)
  // with stage bit 1 -- consecutive lanes then walk one node row across the x-neighbouring elements of the batch
  string column_decode(const string &t, const string &x) const {
    if (!ilv) return "const int ij" + x + " = " + t + " % " + S(P2) + ", le" + x + " = " + t + " / " + S(P2) + ";";
    return "const int le" + x + " = (" + t + " / " + S(P) + ") % " + S(E) + ", ij" + x + " = " + t + " % " + S(P) + " + (" + t + " / " + S(P * E) + ") * " + S(P) + ";";
  }

  // ---- gather + Z ----------------------------------------------------------------------------------
  void emit_z() {
    const int  T = E * P2, R = (T + 31) / 32;
    const bool det = plan->scatter_mode == B200_SCATTER_DETERMINISTIC;
    const bool same_table = !det && gin->rstr == gout->rstr;  // atomic scatter through the gather offsets themselves
    const long long cs = gin->rstr->comp_stride;
    c << "// gather + z-contraction; the scatter targets of the batch are loaded here and parked in shared memory\n";
    c << "template <bool TAIL> static __device__ __noinline__ void b200_lean_z(const long long e0" << stage_params() << ") {\n" << stage_locals();
    c << "  double *const smw = B200_SMW;\n  int *const tgs = B200_TGS;\n  const int lane = threadIdx.x & 31;\n  (void)tgs;\n";
    c << "  const int *const ixb = b200a.in_idx[" << gin->slot << "];\n  const int *const txb = b200a.out_idx[" << gout->slot << "];\n";
    c << "  const double *const ub = b200a.in_ptr[" << gin->slot << "];\n  (void)txb;\n";
    if (st_idx) {
      c << "  const int nel = TAIL ? (int)(b200_ne - e0) : " << E << ";\n  (void)nel;\n";
      c << "  const int *const ofs = B200_OFS + B200_SHIFT(ixb + e0 * " << P3 << ", 4);\n";
      c << "  b200_mbar_wait(B200_BAR(0), par);\n";
    }
    if (ilv) {
      // the batch's offsets and targets: coalesced vector loads of the two contiguous table blocks, stored with the padded element pitch
      const int IP = plan->lean_ip, w = ilv_width(), nv = E * P3 / w, R_ld = (nv + 31) / 32;
      const string vt = w == 4 ? "int4" : (w == 2 ? "int2" : "int");
      c << "  const int nel = TAIL ? (int)(lim - e0) : " << E << ";\n  (void)nel;\n";
      c << "  int *const ofs = B200_OFS;\n";
      c << "  {\n    const " << vt << " *const isrc = (const " << vt << " *)(ixb + e0 * " << P3 << ");\n";
      if (!same_table) c << "    const " << vt << " *const tsrc = (const " << vt << " *)(txb + e0 * " << P3 << ");\n";
      c << "    const int nvec = TAIL ? nel * " << P3 / w << " : " << nv << ";\n";
      for (int m = 0; m < R_ld; m++) {
        c << "    const int q" << m << " = lane + " << 32 * m << ";\n";
        c << "    " << vt << " a" << m << " = {}" << (same_table ? "" : ", b" + S(m) + " = {}") << ";\n";
        c << "    if (q" << m << " < nvec) {\n      a" << m << " = __ldg(isrc + q" << m << ");\n";
        if (!same_table) c << "      b" << m << " = __ldg(tsrc + q" << m << ");\n";
        c << "    }\n";
      }
      for (int m = 0; m < R_ld; m++) {
        c << "    if (q" << m << " < nvec) {\n      const int at = (q" << m << " / " << P3 / w << ") * " << IP << " + (q" << m << " % " << P3 / w << ") * " << w << ";\n";
        c << "      *(" << vt << " *)(ofs + at) = a" << m << ";\n";
        if (!same_table) c << "      *(" << vt << " *)(tgs + at) = b" << m << ";\n";
        c << "    }\n";
      }
      c << "  }\n  __syncwarp();\n";
    }
    // rounds are processed in chunks so that at most ~16 index pairs are in flight per lane
    const int chunk = std::max(1, 16 / P);
    for (int r0 = 0; r0 < R; r0 += chunk) {
      const int r1 = std::min(R, r0 + chunk);
      c << "  {\n";
      for (int r = r0; r < r1; r++) {
        const string x = "_" + S(r);
        c << "    const int t" << x << " = lane + " << 32 * r << ", tc" << x << " = " << ((r + 1) * 32 > T ? "t" + x + " < " + S(T) + " ? t" + x + " : " + S(T - 1) : "t" + x)
          << ";\n";
        c << "    " << column_decode("tc" + x, x) << "\n";
        if (st_idx || ilv) {
          c << "    const int *const ix" << x << " = ofs + (TAIL ? (le" << x << " < nel ? le" << x << " : nel - 1) : le" << x << ") * " << (ilv ? plan->lean_ip : P3) << " + ij" << x
            << ";\n";
          for (int k = 0; k < P; k++) c << "    const int o" << k << x << " = ix" << x << "[" << k * P2 << "];\n";
        } else {
          c << "    const long long e" << x << " = " << elem("le" + x) << ";\n";
          c << "    const int *const ix" << x << " = ixb + e" << x << " * " << P3 << " + ij" << x << ";\n";
          for (int k = 0; k < P; k++) c << "    const int o" << k << x << " = __ldg(ix" << x << " + " << k * P2 << ");\n";
        }
      }
      if (!same_table && !st_idx && !ilv) {
        for (int r = r0; r < r1; r++) {
          const string x = "_" + S(r);
          c << "    const int *const tx" << x << " = txb + e" << x << " * " << P3 << " + ij" << x << ";\n";
          for (int k = 0; k < P; k++) c << "    const int g" << k << x << " = __ldg(tx" << x << " + " << k * P2 << ");\n";
        }
      }
      for (int r = r0; r < r1; r++) {
        const string x = "_" + S(r);
        for (int cc = 0; cc < nc; cc++)
          for (int k = 0; k < P; k++) c << "    const double u" << cc << "_" << k << x << " = __ldg(ub + o" << k << x << " + " << cc * cs << "LL);\n";
      }
      for (int r = r0; r < r1 && !st_idx && !ilv; r++) {
        const string x = "_" + S(r), g = same_table ? "o" : "g";
        int k = 0;
        while (k < P) {
          const int    w   = (P - k) >= 4 && k % 4 == 0 && P % 4 == 0 ? 4 : ((P - k) >= 2 && k % 2 == 0 && P % 2 == 0 ? 2 : 1);
          const string idx = "(" + S((long long)(r * P + k) * 32) + " + lane * " + S(w) + ")";
          if (w == 4) c << "    *(int4 *)(tgs + " << idx << ") = make_int4(" << g << k << x << ", " << g << k + 1 << x << ", " << g << k + 2 << x << ", " << g << k + 3 << x << ");\n";
          else if (w == 2) c << "    *(int2 *)(tgs + " << idx << ") = make_int2(" << g << k << x << ", " << g << k + 1 << x << ");\n";
          else c << "    tgs[" << idx << "] = " << g << k << x << ";\n";
          k += w;
        }
      }
      for (int r = r0; r < r1; r++) {
        const string x = "_" + S(r);
        const bool   partial = (r + 1) * 32 > T;
        c << "    " << (partial ? "if (t" + x + " < " + S(T) + ") " : "") << "{\n";
        for (int cc = 0; cc < nc; cc++) {
          c << "      { double *const dst = smw + le" << x << " * " << ES << " + " << cc * PS << " + ij" << x << ";\n";
          for (int q = 0; q < Q; q++) {
            for (int k = 0; k < P; k++) {
              if (k == 0) c << "        double r" << q << " = cB" << bid << "[" << q * P << "] * u" << cc << "_0" << x << ";\n";
              else c << "        r" << q << " = fma(cB" << bid << "[" << q * P + k << "], u" << cc << "_" << k << x << ", r" << q << ");\n";
            }
          }
          for (int q = 0; q < Q; q++) c << "        dst[" << q * Q * P << "] = r" << q << ";\n";
          c << "      }\n";
        }
        c << "    }\n";
      }
      c << "  }\n";
    }
    c << "}\n\n";
  }

  // ---- streamed inputs of one x-line round ----------------------------------------------------------
  void emit_x_decode(int r, int T) {
    const string x = "_" + S(r);
    c << "  const int t" << x << " = lane + " << 32 * r << ", tc" << x << " = " << ((r + 1) * 32 > T ? "t" + x + " < " + S(T) + " ? t" + x + " : " + S(T - 1) : "t" + x) << ";\n";
    c << "  const int row" << x << " = tc" << x << " % " << Q * Q << ", le" << x << " = tc" << x << " / " << Q * Q << ";\n";
    c << "  const long long e" << x << " = " << elem("le" + x) << ";\n  (void)row" << x << "; (void)e" << x << ";\n";
  }
  void emit_x_loads(int r, bool want_staged) {
    const string x = "_" + S(r);
    const int    Q3 = Q * Q * Q;
    for (size_t f = 0; f < plan->in_fields.size(); f++) {
      const B200GenField &fd = plan->in_fields[f];
      if (fd.emode != B200_EVAL_NONE) continue;
      if (qd_staged(fd) != want_staged) continue;
      const B200Restriction rs = fd.rstr;
      if (want_staged) {
        c << "  const int lq" << f << x << " = (TAIL ? (le" << x << " < nel ? le" << x << " : nel - 1) : le" << x << ") * " << Q3 << " + row" << x << " * " << Q << ";\n";
        for (int cc = 0; cc < fd.nc; cc++)
          for (int q = 0; q < Q; q++) c << "  const double d" << f << "_" << cc * Q + q << x << " = qs" << f << "_" << cc << "[lq" << f << x << " + " << q << "];\n";
        continue;
      }
      c << "  const double *const qd" << f << x << " = b200a.in_ptr[" << fd.slot << "] + e" << x << " * " << rs->strides[2] << "LL + (long long)row" << x << " * "
        << (long long)Q * rs->strides[0] << "LL;\n";
      if (qd_vector(fd)) {
        // 16-byte loads from the aligned window [wp, wp + NV): the fast path needs the window inside the array (the first / last line of
        // the array may stick out by one double) and, for an even Q, an aligned line; everything else takes scalar loads into the same slots
        const int NV = qd_nv();
        c << "  const unsigned long long qlo" << f << x << " = (unsigned long long)b200a.in_ptr[" << fd.slot << "], qhi" << f << x << " = qlo" << f << x << " + "
          << (long long)rs->l_size * 8 << "ULL;\n";
        for (int cc = 0; cc < fd.nc; cc++) {
          const string t = S((long long)f) + "_" + S(cc), w = "w" + t + "_";
          c << "  const unsigned long long qa" << t << x << " = (unsigned long long)(qd" << f << x << " + " << (long long)cc * rs->strides[1] << "LL);\n";
          c << "  const bool od" << t << x << " = " << (Q % 2 ? "((qa" + t + x + " >> 3) & 1ULL) != 0" : "false") << ";\n  (void)od" << t << x << ";\n";
          c << "  double2 ";
          for (int j = 0; j < NV; j++) c << (j ? ", " : "") << w << j << x;
          c << ";\n";
          c << "  { const double2 *const wp = (const double2 *)(qa" << t << x << " & ~15ULL);\n";
          c << "    if ((unsigned long long)wp >= qlo" << f << x << " && (unsigned long long)(wp + " << NV << ") <= qhi" << f << x
            << (Q % 2 ? "" : " && (qa" + t + x + " & 15ULL) == 0") << ") {\n";
          for (int j = 0; j < NV; j++) c << "      " << w << j << x << " = __ldg(wp + " << j << ");\n";
          c << "    } else {\n      const double *const sp = (const double *)qa" << t << x << ";\n";
          for (int q = 0; q < Q; q++) c << "      const double s" << q << " = __ldg(sp + " << q << ");\n";
          // slot j of the window holds value j (even line) or value j - 1 (odd line)
          for (int j = 0; j < 2 * NV; j++) {
            const string ev = j < Q ? "s" + S(j) : string("0.0"), ov = (j >= 1 && j - 1 < Q) ? "s" + S(j - 1) : string("0.0");
            c << "      " << w << j / 2 << x << (j % 2 ? ".y" : ".x") << " = " << (Q % 2 ? "od" + t + x + " ? " + ov + " : " + ev : ev) << ";\n";
          }
          c << "    }\n  }\n";
        }
        continue;
      }
      for (int cc = 0; cc < fd.nc; cc++)
        for (int q = 0; q < Q; q++)
          c << "  const double d" << f << "_" << cc * Q + q << x << " = __ldg(qd" << f << x << " + " << (long long)cc * rs->strides[1] + (long long)q * rs->strides[0] << "LL);\n";
    }
  }

  // ---- Y (in place) + x-lines -------------------------------------------------------------------------
  void emit_yx() {
    B200QFunction qf = op->qf;
    const int     TX = E * Q * Q, RX = (TX + 31) / 32;
    const int     TY = E * nc * P * Q, RY = (TY + 31) / 32;
    c << "// y-contraction in place, then x-lines: X, QFunction on the Q points of the line, X^T (all components of a line per lane)\n";
    c << "template <bool TAIL> static __device__ __noinline__ void b200_lean_yx(const long long e0" << stage_params() << ") {\n" << stage_locals();
    c << "  double *const smw = B200_SMW;\n  const int lane = threadIdx.x & 31;\n";
    bool any_staged_qd = false;
    for (auto &fd : plan->in_fields) any_staged_qd = any_staged_qd || qd_staged(fd);
    if (staged) c << "  const int nel = TAIL ? (int)(b200_ne - e0) : " << E << ";\n  (void)nel;\n";
    // streamed inputs of the first x-line round: in flight across the y-contraction
    emit_x_decode(0, TX);
    emit_x_loads(0, false);
    for (int r = 0; r < RY; r++) {
      const bool partial = (r + 1) * 32 > TY;
      c << "  { const int t = lane + " << 32 * r << ";\n";
      c << "    " << (partial ? "if (t < " + S(TY) + ") " : "") << "{\n";
      c << "      double *const ln = smw + (t / " << P << ") * " << Q * P << " + (t % " << P << ")" << (ES != nc * PS ? " + (t / " + S(nc * P * Q) + ") * " + S(ES - nc * PS) : "")
        << ";\n";
      for (int k = 0; k < P; k++) c << "      const double u" << k << " = ln[" << k * P << "];\n";
      contract("cB" + S(bid), P, Q, false, "u", "r", "      ");
      for (int q = 0; q < Q; q++) c << "      ln[" << q * P << "] = r" << q << ";\n";
      c << "    }\n  }\n";
    }
    c << "  __syncwarp();\n";
    if (any_staged_qd) {
      // bulk-copied quadrature data of this batch: wait for the copy engine, then plain shared-memory reads
      c << "  b200_mbar_wait(B200_BAR(1), par);\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) {
        const B200GenField &fd = plan->in_fields[f];
        if (!qd_staged(fd)) continue;
        for (int cc = 0; cc < fd.nc; cc++)
          c << "  const double *const qs" << f << "_" << cc << " = (const double *)((const char *)smw + " << fd.qd_off + cc * fd.qd_cs * 8 << ") + B200_SHIFT(b200a.in_ptr["
            << fd.slot << "] + " << (long long)cc * fd.rstr->strides[1] << "LL + e0 * " << Q * Q * Q << "LL, 8);\n";
      }
    }
    for (int r = 0; r < RX; r++) {
      const string x       = "_" + S(r);
      const bool   partial = (r + 1) * 32 > TX;
      emit_x_loads(r, true);
      if (r + 1 < RX) {
        emit_x_decode(r + 1, TX);
        emit_x_loads(r + 1, false);
      }
      c << "  " << (partial ? "if (t" + x + " < " + S(TX) + ") " : "") << "{\n";
      c << "    double *const ln = smw + le" << x << " * " << ES << " + row" << x << " * " << P << ";\n";
      c << "    const CeedScalar *in[" << std::max<size_t>(1, qf->inputs.size()) << "];\n";
      c << "    CeedScalar *out[" << std::max<size_t>(1, qf->outputs.size()) << "];\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) c << "    CeedScalar in_" << f << "[" << plan->in_fields[f].size * Q << "];\n";
      for (size_t f = 0; f < plan->out_fields.size(); f++) c << "    CeedScalar out_" << f << "[" << plan->out_fields[f].size * Q << "];\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) c << "    in[" << f << "] = in_" << f << ";\n";
      for (size_t f = 0; f < plan->out_fields.size(); f++) c << "    out[" << f << "] = out_" << f << ";\n";
      for (size_t f = 0; f < plan->in_fields.size(); f++) {
        const B200GenField &fd = plan->in_fields[f];
        if (fd.emode == B200_EVAL_NONE) {
          for (int i = 0; i < fd.nc * Q; i++) c << "    in_" << f << "[" << i << "] = " << qd_value(f, i / Q, i % Q, x) << ";\n";
        } else if (fd.emode == B200_EVAL_WEIGHT) {
          c << "    { const double wyz = cW" << fd.basis_id << "[row" << x << " % " << Q << "] * cW" << fd.basis_id << "[row" << x << " / " << Q << "];\n";
          for (int q = 0; q < Q; q++) c << "      in_" << f << "[" << q << "] = cW" << fd.basis_id << "[" << q << "] * wyz;\n";
          c << "    }\n";
        }
      }
      // x-contraction of every component of the line
      for (int cc = 0; cc < nc; cc++) {
        c << "    {\n";
        emit_line_load("ln + " + S(cc * PS), "u", "      ");
        contract("cB" + S(bid), P, Q, false, "u", "r", "      ");
        for (size_t f = 0; f < plan->in_fields.size(); f++) {
          const B200GenField &fd = plan->in_fields[f];
          if (fd.emode != B200_EVAL_INTERP) continue;
          for (int q = 0; q < Q; q++) c << "      in_" << f << "[" << cc * Q + q << "] = r" << q << ";\n";
        }
        c << "    }\n";
      }
      c << "    " << qf->kernel_name << "(b200a.ctx, " << Q << ", in, out);\n";
      for (int cc = 0; cc < nc; cc++) {
        c << "    {\n";
        for (int q = 0; q < Q; q++) {
          string val;
          for (size_t f = 0; f < plan->out_fields.size(); f++)
            if (plan->out_fields[f].emode == B200_EVAL_INTERP) val += (val.empty() ? "" : " + ") + ("out_" + S(f) + "[" + S(cc * Q + q) + "]");
          c << "      const double v" << q << " = " << val << ";\n";
        }
        contract("cB" + S(bid), Q, P, true, "v", "w", "      ");
        emit_line_store("ln + " + S(cc * PS), "w", "      ");
        c << "    }\n";
      }
      c << "  }\n";
    }
    c << "}\n\n";
  }
  // P consecutive doubles of an x-line: 16-byte accesses when P is even (every line then starts at a 16-byte boundary)
  void emit_line_load(const string &ptr, const string &name, const string &ind) {
    if (P % 2 == 0) {
      for (int i = 0; i < P; i += 2) {
        c << ind << "const double2 " << name << "v" << i << " = *(const double2 *)(" << ptr << " + " << i << ");\n";
        c << ind << "const double " << name << i << " = " << name << "v" << i << ".x, " << name << i + 1 << " = " << name << "v" << i << ".y;\n";
      }
    } else {
      for (int i = 0; i < P; i++) c << ind << "const double " << name << i << " = (" << ptr << ")[" << i << "];\n";
    }
  }
  void emit_line_store(const string &ptr, const string &name, const string &ind) {
    if (P % 2 == 0) {
      for (int i = 0; i < P; i += 2) c << ind << "*(double2 *)(" << ptr << " + " << i << ") = make_double2(" << name << i << ", " << name << i + 1 << ");\n";
    } else {
      for (int i = 0; i < P; i++) c << ind << "(" << ptr << ")[" << i << "] = " << name << i << ";\n";
    }
  }

  // ---- Y^T (in place) -----------------------------------------------------------------------------------
  void emit_yt() {
    const int TY = E * nc * P * Q, RY = (TY + 31) / 32;
    c << "// y-contraction^T in place\n";
    c << "static __device__ __noinline__ void b200_lean_yt() {\n";
    c << "  double *const smw = B200_SMW;\n  const int lane = threadIdx.x & 31;\n";
    for (int r = 0; r < RY; r++) {
      const bool partial = (r + 1) * 32 > TY;
      c << "  { const int t = lane + " << 32 * r << ";\n";
      c << "    " << (partial ? "if (t < " + S(TY) + ") " : "") << "{\n";
      c << "      double *const ln = smw + (t / " << P << ") * " << Q * P << " + (t % " << P << ")" << (ES != nc * PS ? " + (t / " + S(nc * P * Q) + ") * " + S(ES - nc * PS) : "")
        << ";\n";
      for (int q = 0; q < Q; q++) c << "      const double u" << q << " = ln[" << q * P << "];\n";
      contract("cB" + S(bid), Q, P, true, "u", "r", "      ");
      for (int k = 0; k < P; k++) c << "      ln[" << k * P << "] = r" << k << ";\n";
      c << "    }\n  }\n";
    }
    c << "}\n\n";
  }

  // ---- Z^T + scatter ---------------------------------------------------------------------------------------
  void emit_zt() {
    const int       T = E * P2, R = (T + 31) / 32;
    const long long cs = gout->rstr->comp_stride;
    const string    sl = S(gout->slot);
    const bool      rmw = runs;  // the tables may carry the read-modify-write bit
    c << "// z-contraction^T + scatter (targets from the per-lane shared-memory slots the gather stage filled)\n";
    c << "template <bool TAIL> static __device__ __noinline__ void b200_lean_zt(const long long e0" << stage_params() << ") {\n" << stage_locals();
    c << "  double *const smw = B200_SMW;\n  const int lane = threadIdx.x & 31;\n";
    if (st_idx) {
      c << "  const int nel = TAIL ? (int)(b200_ne - e0) : " << E << ";\n  (void)nel;\n";
      c << "  const int *const tgs = B200_TGS + B200_SHIFT(b200a.out_idx[" << sl << "] + e0 * " << P3 << ", 4);\n";
      c << "  b200_mbar_wait(B200_BAR(2), par);\n";
    } else if (ilv) {
      const bool same_table = plan->scatter_mode != B200_SCATTER_DETERMINISTIC && gin->rstr == gout->rstr;
      c << "  const int nel = TAIL ? (int)(lim - e0) : " << E << ";\n  (void)nel;\n";
      c << "  const int *const tgs = " << (same_table ? "B200_OFS" : "B200_TGS") << ";  // staged by the gather stage of this batch\n";
    } else {
      c << "  const int *const tgs = B200_TGS;\n";
    }
    c << "  double *const vb = b200a.out_ptr[" << sl << "];\n  double *const hb = b200a.out_aux[" << sl << "];\n  (void)hb;\n";
    c << "  const long long hd = ((long long)hb - (long long)vb) >> 3;  // halo buffer relative to v, in doubles\n  (void)hd;\n";
    for (int r = 0; r < R; r++) {
      const bool partial = (r + 1) * 32 > T;
      c << "  {\n";
      c << "    const int t = lane + " << 32 * r << ", tc = " << (partial ? "t < " + S(T) + " ? t : " + S(T - 1) : string("t")) << ";\n";
      c << "    " << column_decode("tc", "") << "\n";
      if (st_idx || ilv) {
        c << "    const int *const tx = tgs + (TAIL ? (le < nel ? le : nel - 1) : le) * " << (ilv ? plan->lean_ip : P3) << " + ij;\n";
        for (int k = 0; k < P; k++) c << "    const int g" << k << " = tx[" << k * P2 << "];\n";
      } else {
        int k = 0;
        while (k < P) {
          const int    w   = (P - k) >= 4 && k % 4 == 0 && P % 4 == 0 ? 4 : ((P - k) >= 2 && k % 2 == 0 && P % 2 == 0 ? 2 : 1);
          const string idx = "(" + S((long long)(r * P + k) * 32) + " + lane * " + S(w) + ")";
          if (w == 4) {
            c << "    const int4 gv" << k << " = *(const int4 *)(tgs + " << idx << ");\n";
            c << "    const int g" << k << " = gv" << k << ".x, g" << k + 1 << " = gv" << k << ".y, g" << k + 2 << " = gv" << k << ".z, g" << k + 3 << " = gv" << k << ".w;\n";
          } else if (w == 2) {
            c << "    const int2 gv" << k << " = *(const int2 *)(tgs + " << idx << ");\n";
            c << "    const int g" << k << " = gv" << k << ".x, g" << k + 1 << " = gv" << k << ".y;\n";
          } else {
            c << "    const int g" << k << " = tgs[" << idx << "];\n";
          }
          k += w;
        }
      }
      c << "    const bool live = " << (partial ? "(t < " + S(T) + ")" : string("true")) << " && (!TAIL || e0 + le * st < lim);\n";
      const bool det = plan->scatter_mode != B200_SCATTER_ATOMIC;
      const bool one_store = det && !rmw && !add;  // owner entries and halo entries through ONE store: halo slots addressed relative to v
      if (one_store) {
        for (int k = 0; k < P; k++)
          c << "    const long long a" << k << " = (long long)(g" << k << " ^ (g" << k << " >> 31)) + (g" << k << " < 0 ? hd : 0LL);  // g >= 0: L-index; g < 0: ~g = halo slot\n";
      } else if (det) {
        // direct entries: bit 30 = read-modify-write of a value this warp stored in an earlier iteration (run scatter); the old values are
        // requested (L2, ld.cg) before the contraction so that their latency overlaps it
        for (int k = 0; k < P; k++) {
          if (rmw) c << "    const bool m" << k << " = g" << k << " >= 0 && (" << (add ? "true" : "(g" + S(k) + " & " + S(B200_RUN_RMW_BIT) + ") != 0") << ");\n";
          else c << "    const bool m" << k << " = g" << k << " >= 0 && " << (add ? "true" : "false") << ";\n";
          c << "    const long long a" << k << " = g" << k << (rmw ? " & " + S(B200_RUN_RMW_BIT - 1) : "") << ";\n";
        }
        if (rmw || add)
          for (int cc = 0; cc < nc; cc++)
            for (int k = 0; k < P; k++) c << "    const double old" << cc << "_" << k << " = m" << k << " ? __ldcg(vb + a" << k << " + " << cc * cs << "LL) : 0.0;\n";
      }
      for (int cc = 0; cc < nc; cc++) {
        c << "    {\n      const double *const src = smw + le * " << ES << " + " << cc * PS << " + ij;\n";
        for (int q = 0; q < Q; q++) c << "      const double u" << q << " = src[" << q * Q * P << "];\n";
        contract("cB" + S(bid), Q, P, true, "u", "r", "      ");
        c << "      if (live) {\n";
        for (int k = 0; k < P; k++) {
          if (!det) {
            c << "        atomicAdd(vb + g" << k << " + " << cc * cs << "LL, r" << k << ");\n";
          } else if (one_store) {
            if (cc == 0) c << "        vb[a" << k << "] = r" << k << ";\n";
            else c << "        vb[a" << k << " + (g" << k << " < 0 ? " << cc << " * b200a.ord_num_halo : " << cc * cs << "LL)] = r" << k << ";\n";
          } else {
            c << "        if (g" << k << " >= 0) vb[a" << k << " + " << cc * cs << "LL] = " << ((rmw || add) ? "m" + S(k) + " ? old" + S(cc) + "_" + S(k) + " + r" + S(k) + " : r" + S(k) : "r" + S(k)) << ";\n";
            c << "        else hb[(long long)(~g" << k << ") + " << cc << " * b200a.ord_num_halo] = r" << k << ";\n";
          }
        }
        c << "      }\n    }\n";
      }
      c << "  }\n";
    }
    c << "}\n\n";
  }

  // ---- in-kernel finalize (stage bit 128) ------------------------------------------------------------------
  // part table b200a.ord_pred_ptr: [0] = K, [1 + c] = first batch of part c (c = 0..K), [K + 2 + c] = number of shared nodes completed by
  // parts < c (c = 0..K); b200a.ord_flags = per-part counters of finished batches (zeroed before the launch); b200a.ord_pred_idx /
  // b200a.ord_sync = halo_node / halo_ptr of the output restriction.
  void emit_fin() {
    const long long cs = gout->rstr->comp_stride;
    const string    sl = S(gout->slot);
    c << "// this warp has finished `mine` batches of part c: publish (release) -- its stores to v and to the halo buffer come first\n";
    c << "static __device__ __forceinline__ void b200_fin_signal(const int c, const int mine) {\n  __syncwarp();\n"
      << "  if ((threadIdx.x & 31) == 0 && mine) {\n    __threadfence();\n    atomicAdd(b200a.ord_flags + c, mine);\n  }\n}\n";
    c << "// fold this warp's slice of the shared nodes completed by part c into v, once every batch of the part has been published\n";
    c << "static __device__ __noinline__ void b200_fin_part(const int c, const long long gw, const long long ng) {\n";
    c << "  const int *const tab = b200a.ord_pred_ptr;\n  const int K = tab[0], total = tab[2 + c] - tab[1 + c];\n  const int lane = threadIdx.x & 31;\n";
    c << "  if (lane == 0) {\n    int seen;\n    do { asm volatile(\"ld.acquire.gpu.global.s32 %0, [%1];\" : \"=r\"(seen) : \"l\"(b200a.ord_flags + c) : \"memory\"); } while (seen < total);\n  }\n";
    c << "  __syncwarp();\n";
    c << "  const long long s0 = tab[K + 2 + c], cnt = tab[K + 3 + c] - s0;\n";
    c << "  const long long b = s0 + gw * cnt / ng, e = s0 + (gw + 1) * cnt / ng;\n";
    c << "  const int *const hn = b200a.ord_pred_idx;\n  const int *const hp = b200a.ord_sync;\n";
    c << "  const double *const h = b200a.out_aux[" << sl << "];\n  double *const v = b200a.out_ptr[" << sl << "];\n";
    // U nodes per lane in flight: the table loads of all of them first, then their owner values and first halo values, then the (short)
    // tails -- the standalone finalize kernel gets its memory parallelism from 1.2 M threads, a persistent grid has to unroll for it
    const int U = 8;
    c << "  for (long long i0 = b + lane; i0 < e; i0 += " << 32 * U << ") {\n";
    c << "    long long l[" << U << "];\n    int p0[" << U << "], p1[" << U << "];\n";
    c << "#pragma unroll\n    for (int u = 0; u < " << U << "; u++) {\n      const long long i = i0 + 32 * u;\n      const bool ok = i < e;\n"
      << "      l[u] = ok ? (long long)__ldg(hn + i) : -1LL;\n      p0[u] = ok ? __ldg(hp + i) : 0;\n      p1[u] = ok ? __ldg(hp + i + 1) : 0;\n    }\n";
    for (int cc = 0; cc < nc; cc++) {
      c << "    {\n      double acc[" << U << "], h0[" << U << "];\n";
      c << "#pragma unroll\n      for (int u = 0; u < " << U << "; u++) acc[u] = l[u] >= 0 ? __ldcg(v + l[u] + " << cc * cs << "LL) : 0.0;\n";
      c << "#pragma unroll\n      for (int u = 0; u < " << U << "; u++) h0[u] = l[u] >= 0 ? __ldcg(h + p0[u] + " << cc << " * b200a.ord_num_halo) : 0.0;\n";
      c << "#pragma unroll\n      for (int u = 0; u < " << U << "; u++) {\n        double a = acc[u] + h0[u];\n"
        << "        for (int j = p0[u] + 1; j < p1[u]; j++) a += __ldcg(h + j + " << cc << " * b200a.ord_num_halo);\n"
        << "        if (l[u] >= 0) v[l[u] + " << cc * cs << "LL] = a;\n      }\n    }\n";
    }
    c << "  }\n}\n\n";
  }

  string generate() {
    gin  = &plan->in_groups[0];
    gout = &plan->out_groups[0];
    bid  = gin->basis_id;
    P    = plan->bases[bid].P;
    Q    = plan->Q;
    E    = plan->epb;
    nc   = gin->nc;
    P2   = P * P;
    P3   = P2 * P;
    PS   = Q * Q * P;
    ES   = plan->lean_es;
    ilv    = (plan->stage_mask & 1) != 0;
    vqd    = (plan->stage_mask & 4) != 0;
    st_idx = (plan->stage_mask & 8) != 0;
    st_qd  = (plan->stage_mask & 32) != 0;
    staged = st_idx || st_qd;
    pf_qd  = (plan->stage_mask & 64) != 0 && !st_qd;
    fin    = (plan->stage_mask & 128) != 0 && !staged && !runs && plan->scatter_mode == B200_SCATTER_DETERMINISTIC;
    runs   = plan->lean_runs && !staged;
    emit_header();
    emit_z();
    emit_yx();
    emit_yt();
    emit_zt();
    if (fin) emit_fin();
    const int NT = plan->threads, W = NT / 32;
    c << "extern \"C\" __global__ void __launch_bounds__(" << NT << ", " << std::max(1, plan->blocks_per_sm) << ") b200_operator_" << op->qf->kernel_name
      << "(const long long e_begin, const long long e_end, const int run_mode) {\n";
    c << "  if (threadIdx.x == 0) b200_ne = e_end;\n  __syncthreads();\n";
    if (staged) c << "  const long long num_batches = (e_end - e_begin + " << E - 1 << ") / " << E << ";\n  (void)run_mode;\n";
    const string first = "(long long)blockIdx.x * " + S(W) + " + (threadIdx.x >> 5)", stride = "(long long)gridDim.x * " + S(W);
    if (!staged) {
      const string a = stage_args();
      if (runs) {
        // run mode (run_mode != 0, deterministic scatter with run tables): the warp owns one contiguous run of elements and walks it
        // in nb iterations, iteration i holding elements s + i + k * nb; otherwise grid-stride over contiguous batches of E elements
        c << "  const long long ne = e_end - e_begin, gw = " << first << ", ng = " << stride << ";\n";
        c << "  long long base0, step, lim;\n  int st, n_it;\n";
        c << "  if (run_mode) {\n    base0 = e_begin + gw * ne / ng;\n    lim = e_begin + (gw + 1) * ne / ng;\n"
          << "    n_it = (int)((lim - base0 + " << E - 1 << ") / " << E << ");\n    st = n_it;\n    step = 1;\n  } else {\n"
          << "    const long long num_batches = (ne + " << E - 1 << ") / " << E << ";\n    n_it = gw < num_batches ? (int)((num_batches - gw + ng - 1) / ng) : 0;\n"
          << "    base0 = e_begin + gw * " << E << ";\n    step = ng * " << E << ";\n    st = 1;\n    lim = e_end;\n  }\n";
        c << "  for (int it = 0; it < n_it; it++) {\n";
        c << "    const long long e0 = base0 + it * step;\n";
        c << "    if (e0 + " << E - 1 << "LL * st < lim) {\n";
      } else {
        c << "  (void)run_mode;\n  const long long num_batches = (e_end - e_begin + " << E - 1 << ") / " << E << ";\n";
        if (fin) {
          // run_mode == 2: in-kernel finalize (whole-mesh launch, e_begin == 0); parts are ranges of the global batch index
          c << "  const long long gw = " << first << ", ng = " << stride << ";\n";
          c << "  const int *const tab = b200a.ord_pred_ptr;\n  const int K = run_mode == 2 ? tab[0] : 0;\n  int part = 0, mine = 0;\n";
        }
        c << "  for (long long batch = " << first << "; batch < num_batches; batch += " << stride << ") {\n";
        if (fin)
          c << "    while (part < K && batch >= tab[2 + part]) {  // this warp is done with part `part`\n      b200_fin_signal(part, mine);\n      mine = 0;\n"
            << "      if (part >= 1) b200_fin_part(part - 1, gw, ng);\n      part++;\n    }\n    mine++;\n";
        c << "    const long long e0 = e_begin + batch * " << E << ";\n";
        if (pf_qd)
          c << "    { const long long e0n = e_begin + (batch + " << stride << ") * " << E << ";\n"
            << "      if ((threadIdx.x & 31) == 0 && e0n < e_end) b200_lean_prefetch_qd(e0n, (int)(e_end - e0n < " << E << " ? e_end - e0n : " << E << ")); }\n";
        c << "    if (e0 + " << E << " <= e_end) {\n";
      }
      c << "      b200_lean_z<false>(e0" << a << ");\n      __syncwarp();\n      b200_lean_yx<false>(e0" << a << ");\n      __syncwarp();\n      b200_lean_yt();\n      __syncwarp();\n"
        << "      b200_lean_zt<false>(e0" << a << ");\n      __syncwarp();\n";
      c << "    } else {\n";
      c << "      b200_lean_z<true>(e0" << a << ");\n      __syncwarp();\n      b200_lean_yx<true>(e0" << a << ");\n      __syncwarp();\n      b200_lean_yt();\n      __syncwarp();\n"
        << "      b200_lean_zt<true>(e0" << a << ");\n      __syncwarp();\n";
      c << "    }\n  }\n";
      if (fin && !runs)
        c << "  for (; part < K; part++) {\n    b200_fin_signal(part, mine);\n    mine = 0;\n    if (part >= 1) b200_fin_part(part - 1, gw, ng);\n  }\n"
          << "  if (K) b200_fin_part(K - 1, gw, ng);\n";
      c << "}\n";
      return c.str();
    }
    bool any_qd = false;
    for (auto &fd : plan->in_fields) any_qd = any_qd || qd_staged(fd);
    // bulk pipeline: every stream buffer is re-armed for the warp's next batch as soon as its consumer stage is done
    c << "  const bool lead = (threadIdx.x & 31) == 0;\n";
    c << "  if (lead) { b200_mbar_init(B200_BAR(0), 1); b200_mbar_init(B200_BAR(1), 1); b200_mbar_init(B200_BAR(2), 1); }\n  __syncwarp();\n";
    c << "  long long batch = " << first << ";\n";
    c << "  if (lead && batch < num_batches) {\n    const long long e0 = e_begin + batch * " << E << ";\n"
      << "    const int ne = (int)(e_end - e0 < " << E << " ? e_end - e0 : " << E << ");\n"
      << (st_idx ? "    b200_lean_issue_off(e0, ne);\n" : "") << (any_qd ? "    b200_lean_issue_qd(e0, ne);\n" : "") << (st_idx ? "    b200_lean_issue_tg(e0, ne);\n" : "") << "  }\n"
      << "  __syncwarp();  // (trailing elements the issuing lane moved itself at the end of an array)\n";
    c << "  for (int it = 0; batch < num_batches; batch += " << stride << ", it++) {\n";
    c << "    const long long e0 = e_begin + batch * " << E << ", e0n = e_begin + (batch + " << stride << ") * " << E << ";\n";
    c << "    const int nen = e0n >= e_end ? 0 : (int)(e_end - e0n < " << E << " ? e_end - e0n : " << E << "), par = it & 1;\n";
    if (pf_qd) c << "    if (lead && nen) b200_lean_prefetch_qd(e0n, nen);\n";
    for (int tail = 0; tail < 2; tail++) {
      const string T = tail ? "true" : "false";
      c << (tail ? "    } else {\n" : "    if (e0 + " + S(E) + " <= e_end) {\n");
      c << "      b200_lean_z<" << T << ">(e0, par);\n      __syncwarp();\n" << (st_idx ? "      if (lead && nen) b200_lean_issue_off(e0n, nen);\n" : "");
      c << "      b200_lean_yx<" << T << ">(e0, par);\n      __syncwarp();\n" << (any_qd ? "      if (lead && nen) b200_lean_issue_qd(e0n, nen);\n" : "");
      c << "      b200_lean_yt();\n      __syncwarp();\n";
      c << "      b200_lean_zt<" << T << ">(e0, par);\n      __syncwarp();\n" << (st_idx ? "      if (lead && nen) b200_lean_issue_tg(e0n, nen);\n" : "");
    }
    c << "    }\n  }\n}\n";
    return c.str();
  }
};

}  // namespace

// Can this plan (fields already classified by b200_opgen_plan) run on the lean kernel?
bool b200_opgen_lean_eligible(const B200OpPlan *plan) {
  if (plan->in_groups.size() != 1 || plan->out_groups.size() != 1) return false;
  const B200GenGroup &gi = plan->in_groups[0], &go = plan->out_groups[0];
  if (gi.basis_id != go.basis_id || gi.nc != go.nc || gi.use_grad || go.use_grad) return false;
  if (gi.rstr->is_strided || go.rstr->is_strided) return false;
  const B200GenBasis &b = plan->bases[gi.basis_id];
  if (b.collocated || b.P > b.Q || b.P > 10 || plan->Q > 12) return false;
  if (plan->scatter_mode != B200_SCATTER_DETERMINISTIC && plan->scatter_mode != B200_SCATTER_ATOMIC) return false;
  for (auto &f : plan->in_fields) {
    if (f.emode == B200_EVAL_NONE && !f.rstr->is_strided) return false;
    if (f.emode != B200_EVAL_NONE && f.emode != B200_EVAL_WEIGHT && f.emode != B200_EVAL_INTERP) return false;
  }
  for (auto &f : plan->out_fields)
    if (f.emode != B200_EVAL_INTERP) return false;
  return true;
}

// Shared memory of one warp for E elements per iteration: the in-place planes, then the parked scatter targets.
size_t b200_opgen_lean_layout(B200OpPlan *plan, int E) {
  const B200GenGroup &gi = plan->in_groups[0];
  const int           P  = plan->bases[gi.basis_id].P, Q = plan->Q;
  plan->lean_es          = gi.nc * Q * Q * P;
  // stage bit 2: element stride = P (mod 16) doubles, so that the lanes (i, element) of an interleaved column stage hit consecutive banks
  // (an even P keeps an even stride: x-lines stay 16-byte aligned)
  if (plan->stage_mask & 2)
    while (plan->lean_es % 16 != P % 16) plan->lean_es++;
  size_t off             = ((size_t)E * plan->lean_es * 8 + 15) / 16 * 16;
  auto   take            = [&](size_t bytes) {
    const size_t at = off;
    off += (bytes + 15) / 16 * 16;
    return (int)at;
  };
  plan->mbar_off = -1, plan->lean_off_off = -1;
  for (auto &f : plan->in_fields) f.qd_off = -1, f.qd_tma = false;
  // bulk pipeline: single buffers for the offsets and the targets (bit 8) and every contiguous EVAL_NONE component (bit 32), each
  // + 16 bytes for the aligned-down sources
  if (plan->stage_mask & 1) {
    // interleaved columns: both tables of the batch staged with an element pitch = P (mod 32) ints
    int ip = P * P * P;
    while (ip % 32 != P % 32) ip++;
    plan->lean_ip      = ip;
    plan->lean_off_off = take((size_t)E * ip * 4);
    plan->lean_tg_off  = take((size_t)E * ip * 4);
  } else if (plan->stage_mask & 8) {
    plan->lean_off_off = take((size_t)E * P * P * P * 4 + 16);
    plan->lean_tg_off  = take((size_t)E * P * P * P * 4 + 16);
  } else {
    const int rounds  = (E * P * P + 31) / 32;
    plan->lean_tg_off = take((size_t)rounds * P * 32 * 4);
  }
  if (plan->stage_mask & 32) {
    for (auto &f : plan->in_fields) {
      const bool contiguous = f.emode == B200_EVAL_NONE && f.rstr->is_strided && f.rstr->strides[0] == 1 && f.rstr->strides[2] == f.rstr->elem_size;
      if (!contiguous) continue;
      f.qd_cs  = E * Q * Q * Q + 2;
      f.qd_off = take((size_t)f.nc * f.qd_cs * 8);
      f.qd_tma = true;
    }
  }
  if (plan->stage_mask & 40) plan->mbar_off = take(32);
  return off;
}

std::string b200_opgen_lean_source(B200Operator op, B200OpPlan *plan, int add) {
  LeanGen g;
  g.op   = op;
  g.plan = plan;
  g.add  = add;
  return g.generate();
}
