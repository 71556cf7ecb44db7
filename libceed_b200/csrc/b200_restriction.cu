// b200_restriction.cu -- CeedElemRestriction for the b200 backend (standard offsets + strided).
//
// Behaviour follows backends/cuda-ref/ceed-cuda-ref-restriction.c:
//   * E-vector layout {1, num_elem*elem_size, elem_size} = [comp][elem][node]           (:530-541)
//   * NoTranspose:  v[n + e*S + c*S*Ne] = u[offsets[n + e*S] + c*comp_stride]            (cuda-ref-restriction-offset.h:15-26)
//   * Transpose is DETERMINISTIC: per referenced L-node, E-entries are summed in ascending (elem,node) order -- the
//     order of the serial CPU loop in backends/ref/ceed-ref-restriction.c:220-242 -- through a transpose CSR built
//     at setup (cuda-ref-restriction.c:416-493), so results are bit-identical to /cpu/self/ref/serial.
//   * strided: L-index = n*strides[0] + c*strides[1] + e*strides[2]                      (cuda-ref-restriction-strided.h:15-39)
// In addition, an "owner/halo" decomposition of the transpose is built for the fused operator kernel (see
// b200_restriction_build_owner): the first E-entry of every L-node owns the store, the remaining entries go through
// a compact halo buffer that a finalize kernel folds in, again in ascending E-order.
// Index arithmetic on E-/Q-sized ranges in the kernels is 64-bit (CeedInt overflows at 100M DoFs); the setup
// tables (CSR rows, scatter targets, halo slots) hold E-entry indices in int32 and refuse restrictions with >= 2^31 - 1 entries.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "b200_internal.h"

namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(B200Ceed ceed, int64_t n) {
  int64_t blocks = (n + kThreads - 1) / kThreads;
  int64_t cap    = (int64_t)ceed->num_sms * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

__global__ void k_offset_notranspose(const int32_t *__restrict__ offsets, const double *__restrict__ u, double *__restrict__ v, int64_t e_entries,
                                     int num_comp, int64_t comp_stride) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_entries; i += stride) {
    const int64_t l = offsets[i];
    for (int c = 0; c < num_comp; c++) v[i + c * e_entries] = u[l + c * comp_stride];
  }
}

// One thread per referenced L-node; fixed summation order => deterministic and equal to the serial reference.
__global__ void k_offset_transpose(const int32_t *__restrict__ lvec_indices, const int32_t *__restrict__ t_offsets,
                                   const int32_t *__restrict__ t_indices, const double *__restrict__ u, double *__restrict__ v, int64_t num_nodes,
                                   int64_t e_entries, int num_comp, int64_t comp_stride) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_nodes; i += stride) {
    const int64_t l     = lvec_indices[i];
    const int32_t begin = t_offsets[i], end = t_offsets[i + 1];
    for (int c = 0; c < num_comp; c++) {
      double acc = v[l + c * comp_stride];
      for (int32_t j = begin; j < end; j++) acc += u[(int64_t)t_indices[j] + c * e_entries];
      v[l + c * comp_stride] = acc;
    }
  }
}

__global__ void k_strided_notranspose(const double *__restrict__ u, double *__restrict__ v, int64_t num_elem, int elem_size, int num_comp, int64_t s0,
                                      int64_t s1, int64_t s2) {
  int64_t e_entries = num_elem * elem_size;
  int64_t stride    = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_entries; i += stride) {
    const int64_t e = i / elem_size, n = i - e * elem_size;
    for (int c = 0; c < num_comp; c++) v[i + c * e_entries] = u[n * s0 + c * s1 + e * s2];
  }
}
__global__ void k_strided_transpose(const double *__restrict__ u, double *__restrict__ v, int64_t num_elem, int elem_size, int num_comp, int64_t s0,
                                    int64_t s1, int64_t s2) {
  int64_t e_entries = num_elem * elem_size;
  int64_t stride    = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_entries; i += stride) {
    const int64_t e = i / elem_size, n = i - e * elem_size;
    for (int c = 0; c < num_comp; c++) v[n * s0 + c * s1 + e * s2] += u[i + c * e_entries];
  }
}

// Fold the halo buffer of the fused operator kernel into v: v[node] += sum of the node's non-owner contributions
// (ascending E-order).  halo is [comp][num_halo].
__global__ void k_halo_finalize(const int32_t *__restrict__ halo_node, const int32_t *__restrict__ halo_ptr, const double *__restrict__ halo,
                                double *__restrict__ v, int64_t num_shared, int64_t num_halo, int num_comp, int64_t comp_stride) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_shared; i += stride) {
    const int64_t l     = halo_node[i];
    const int32_t begin = halo_ptr[i], end = halo_ptr[i + 1];
    for (int c = 0; c < num_comp; c++) {
      double acc = v[l + c * comp_stride];
      for (int32_t j = begin; j < end; j++) acc += halo[(int64_t)j + c * num_halo];
      v[l + c * comp_stride] = acc;
    }
  }
}

// Dense form of the finalize pass: thread t of block b handles the 4 consecutive L-indices b * 1024 + 4 t ... + 3.  Their halo-slot
// counts come in as one 4-byte load, the first slot of the thread's nodes is the block's base + an exclusive block scan of the counts
// (no per-node pointer table, no node list); same additions in the same order as k_halo_finalize.
constexpr int kDenseThreads = 256, kDenseBlockNodes = 4 * kDenseThreads;
__global__ void __launch_bounds__(kDenseThreads) k_halo_finalize_dense(const uint8_t *__restrict__ cnt, const int32_t *__restrict__ block_base,
                                                                        const double *__restrict__ halo, double *__restrict__ v, int64_t num_halo, int num_comp,
                                                                        int64_t comp_stride) {
  __shared__ int warp_sum[kDenseThreads / 32];
  const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * kDenseBlockNodes + 4 * threadIdx.x;
  const uchar4  c4 = reinterpret_cast<const uchar4 *>(cnt)[n0 >> 2];  // (the table is zero-padded to a whole number of blocks)
  const int     c[4] = {c4.x, c4.y, c4.z, c4.w};
  const int     mine = c[0] + c[1] + c[2] + c[3];
  int           incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int w = 0; w < kDenseThreads / 32; w++) before += w < warp ? warp_sum[w] : 0;
  if (!mine) return;
  int64_t slot = (int64_t)block_base[blockIdx.x] + before + incl - mine;
  for (int comp = 0; comp < num_comp; comp++) {
    const double *h = halo + (int64_t)comp * num_halo;
    double       *w = v + (int64_t)comp * comp_stride + n0;
    int64_t       s = slot;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (c[k]) {
        double acc = w[k];
        for (int j = 0; j < c[k]; j++) acc += h[s + j];
        w[k] = acc;
        s += c[k];
      }
    }
  }
}

}  // namespace

#define LAUNCH(ceed, kernel, n, ...)                                                                                  \
  do {                                                                                                                \
    B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");   \
    kernel<<<grid_for(ceed, n), kThreads, 0, (ceed)->stream>>>(__VA_ARGS__);                                          \
    (ceed)->launch_count++;                                                                                           \
    B200_CUDA(ceed, cudaGetLastError());                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ setup
extern "C" int ceedb200_restriction_create(B200Ceed ceed, b200_int num_elem, b200_int elem_size, b200_int num_comp, b200_int comp_stride,
                                           b200_size l_size, int mem_type, int copy_mode, const b200_int *offsets, B200Restriction *rstr_out) {
  B200_CHECK(num_elem >= 0 && elem_size > 0 && num_comp > 0, ceed, B200_ERROR_DIMENSION, "invalid restriction dimensions");
  B200Restriction r = new B200Restriction_();
  r->ceed           = ceed;
  r->num_elem       = num_elem;
  r->elem_size      = elem_size;
  r->num_comp       = num_comp;
  r->comp_stride    = comp_stride;
  r->l_size         = l_size;
  const int64_t n   = (int64_t)num_elem * elem_size;
  const size_t  bytes = n * sizeof(int32_t);
  if (mem_type == B200_MEM_HOST) {
    switch (copy_mode) {
      case B200_COPY_VALUES:
        r->h_offsets_owned = (int32_t *)malloc(bytes ? bytes : 4);
        memcpy(r->h_offsets_owned, offsets, bytes);
        r->h_offsets = r->h_offsets_owned;
        break;
      case B200_OWN_POINTER:
        r->h_offsets_owned = (int32_t *)offsets;
        r->h_offsets       = offsets;
        break;
      default:
        r->h_offsets_borrowed = offsets;
        r->h_offsets          = offsets;
    }
    B200_CALL(b200_dmalloc(ceed, (void **)&r->d_offsets, bytes));
    r->d_offsets_owned = true;
    B200_CALL(b200_h2d(ceed, r->d_offsets, r->h_offsets, bytes));
  } else {
    // device offsets: keep (copy or adopt) the device array and mirror it on the host for the setup analysis
    if (copy_mode == B200_COPY_VALUES) {
      B200_CALL(b200_dmalloc(ceed, (void **)&r->d_offsets, bytes));
      r->d_offsets_owned = true;
      B200_CALL(b200_d2d(ceed, r->d_offsets, offsets, bytes));
    } else {
      r->d_offsets       = (int32_t *)offsets;
      r->d_offsets_owned = copy_mode == B200_OWN_POINTER;
    }
    r->h_offsets_owned = (int32_t *)malloc(bytes ? bytes : 4);
    B200_CALL(b200_d2h(ceed, r->h_offsets_owned, r->d_offsets, bytes));
    r->h_offsets = r->h_offsets_owned;
  }
  for (int64_t i = 0; i < n; i++) {
    const int64_t l = r->h_offsets[i];
    if (l < 0 || l + (int64_t)(num_comp - 1) * comp_stride >= l_size) {
      int code = b200_error(ceed, B200_ERROR_DIMENSION, "Restriction offset %lld (%lld) out of range [0, %lld)", (long long)i, (long long)l,
                            (long long)l_size);
      ceedb200_restriction_destroy(r);
      return code;
    }
  }
  *rstr_out = r;
  return B200_SUCCESS;
}

extern "C" int ceedb200_restriction_create_strided(B200Ceed ceed, b200_int num_elem, b200_int elem_size, b200_int num_comp, b200_size l_size,
                                                   const b200_int strides[3], B200Restriction *rstr_out) {
  B200_CHECK(num_elem >= 0 && elem_size > 0 && num_comp > 0, ceed, B200_ERROR_DIMENSION, "invalid restriction dimensions");
  B200Restriction r = new B200Restriction_();
  r->ceed           = ceed;
  r->num_elem       = num_elem;
  r->elem_size      = elem_size;
  r->num_comp       = num_comp;
  r->l_size         = l_size;
  r->is_strided     = true;
  if (!strides || (strides[0] == 0 && strides[1] == 0 && strides[2] == 0)) {
    // CEED_STRIDES_BACKEND: same choice as the reference CUDA backends (cuda-gen-operator-build.cpp:466)
    r->backend_strides = true;
    r->strides[0]      = 1;
    r->strides[1]      = (int64_t)elem_size * num_elem;
    r->strides[2]      = elem_size;
  } else {
    for (int i = 0; i < 3; i++) r->strides[i] = strides[i];
  }
  *rstr_out = r;
  return B200_SUCCESS;
}

extern "C" int ceedb200_restriction_destroy(B200Restriction r) {
  if (!r) return B200_SUCCESS;
  B200Ceed ceed = r->ceed;
  free(r->h_offsets_owned);
  if (r->d_offsets_owned) b200_dfree(ceed, r->d_offsets);
  b200_dfree(ceed, r->d_lvec_indices);
  b200_dfree(ceed, r->d_t_offsets);
  b200_dfree(ceed, r->d_t_indices);
  b200_dfree(ceed, r->d_tgt);
  b200_dfree(ceed, r->d_halo_node);
  b200_dfree(ceed, r->d_halo_ptr);
  b200_dfree(ceed, r->d_halo_cnt);
  b200_dfree(ceed, r->d_halo_block_base);
  delete r;
  return B200_SUCCESS;
}

// The setup tables hold E-entry indices, CSR row starts and halo slots in int32.
static int check_table_range(B200Restriction r) {
  B200_CHECK((int64_t)r->num_elem * r->elem_size < INT32_MAX, r->ceed, B200_ERROR_UNSUPPORTED,
             "Backend does not implement restrictions with more than 2^31 - 1 E-vector entries (%lld)", (long long)r->num_elem * r->elem_size);
  return B200_SUCCESS;
}

// Host-side transpose CSR (counting sort by L-index; stable => ascending E-order inside each row).
struct HostTranspose {
  std::vector<int32_t> lvec_indices, t_offsets, t_indices;
};
static void build_host_transpose(B200Restriction r, HostTranspose &t) {
  const int64_t        n = (int64_t)r->num_elem * r->elem_size;
  std::vector<int32_t> count(r->l_size + 1, 0);
  for (int64_t i = 0; i < n; i++) count[r->h_offsets[i] + 1]++;
  int64_t num_nodes = 0;
  for (int64_t l = 0; l < r->l_size; l++) num_nodes += count[l + 1] > 0;
  t.lvec_indices.resize(num_nodes);
  t.t_offsets.resize(num_nodes + 1);
  t.t_indices.resize(n);
  // compact row id per L-index + row starts
  std::vector<int32_t> row_of(r->l_size, -1);
  int64_t              row = 0, pos = 0;
  for (int64_t l = 0; l < r->l_size; l++) {
    if (count[l + 1] > 0) {
      row_of[l]            = (int32_t)row;
      t.lvec_indices[row]  = (int32_t)l;
      t.t_offsets[row]     = (int32_t)pos;
      pos += count[l + 1];
      row++;
    }
  }
  t.t_offsets[num_nodes] = (int32_t)pos;
  std::vector<int32_t> fill(t.t_offsets.begin(), t.t_offsets.end() - 1);
  for (int64_t i = 0; i < n; i++) t.t_indices[fill[row_of[r->h_offsets[i]]]++] = (int32_t)i;
}

int b200_restriction_build_transpose(B200Restriction r) {
  if (r->transpose_built || r->is_strided) return B200_SUCCESS;
  B200Ceed      ceed = r->ceed;
  B200_CALL(check_table_range(r));  // the CSR tables hold E-entry indices in int32
  HostTranspose t;
  build_host_transpose(r, t);
  r->num_nodes = (int64_t)t.lvec_indices.size();
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_lvec_indices, t.lvec_indices.size() * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_t_offsets, t.t_offsets.size() * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_t_indices, t.t_indices.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, r->d_lvec_indices, t.lvec_indices.data(), t.lvec_indices.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, r->d_t_offsets, t.t_offsets.data(), t.t_offsets.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, r->d_t_indices, t.t_indices.data(), t.t_indices.size() * sizeof(int32_t)));
  r->transpose_built = true;
  return B200_SUCCESS;
}

// Owner/halo decomposition used by the fused operator kernel's deterministic scatter:
//   tgt[E-entry] >= 0 : this entry is the FIRST (lowest E-index) reference to its L-node; the kernel stores/accumulates
//                       its value straight into v[tgt] -- exactly one such entry per node, so no race and no atomics.
//   tgt[E-entry] <  0 : ~tgt is a slot of the halo buffer.  Slots are grouped by node (halo_ptr CSR over the compact list
//                       of shared nodes halo_node) and ordered by ascending E-index inside a node, so the finalize kernel
//                       reads them contiguously and adds them in the same order as the serial reference scatter.
int b200_restriction_build_owner(B200Restriction r) {
  if (r->owner_built || r->is_strided) return B200_SUCCESS;
  B200Ceed      ceed = r->ceed;
  B200_CALL(check_table_range(r));
  HostTranspose t;
  build_host_transpose(r, t);
  const int64_t        n         = (int64_t)r->num_elem * r->elem_size;
  const int64_t        num_nodes = (int64_t)t.lvec_indices.size();
  std::vector<int32_t> tgt(n), halo_node, halo_ptr;
  int64_t              slot = 0;
  // shared nodes ordered by the element part of their last toucher (see part_ends; one part when no split is set), L-index order inside
  // a part: the nodes completed by part p can be finalized -- and sent, or copied to the host -- while later parts are still applied
  std::vector<int32_t> ends = r->part_ends;
  if (ends.empty() || ends.back() != r->num_elem) ends.push_back(r->num_elem);
  const int num_parts = (int)ends.size();
  auto part_of = [&](int64_t elem) { return (int)(std::upper_bound(ends.begin(), ends.end(), (int32_t)elem) - ends.begin()); };
  std::vector<int32_t> node_part(num_nodes, -1);
  std::vector<int64_t> count(num_parts + 1, 0);
  for (int64_t row = 0; row < num_nodes; row++) {
    const int32_t begin = t.t_offsets[row], end = t.t_offsets[row + 1];
    tgt[t.t_indices[begin]] = t.lvec_indices[row];
    if (end - begin <= 1) continue;
    node_part[row] = std::min(num_parts - 1, part_of(t.t_indices[end - 1] / r->elem_size));  // last toucher = highest element
    count[node_part[row] + 1]++;
  }
  for (int p = 0; p < num_parts; p++) count[p + 1] += count[p];
  r->shared_prefix.assign(count.begin(), count.end());  // shared_prefix[p] = shared nodes completed by parts < p ... [num_parts] = all
  {
    const int64_t        num_shared = count[num_parts];
    std::vector<int64_t> fill(count.begin(), count.end() - 1);
    std::vector<int32_t> row_at(num_shared);
    for (int64_t row = 0; row < num_nodes; row++)
      if (node_part[row] >= 0) row_at[fill[node_part[row]]++] = (int32_t)row;
    halo_node.resize(num_shared);
    halo_ptr.resize(num_shared);
    for (int64_t i = 0; i < num_shared; i++) {
      const int32_t row   = row_at[i];
      const int32_t begin = t.t_offsets[row], end = t.t_offsets[row + 1];
      halo_node[i]        = t.lvec_indices[row];
      halo_ptr[i]         = (int32_t)slot;
      for (int32_t j = begin + 1; j < end; j++) tgt[t.t_indices[j]] = ~(int32_t)(slot++);
    }
  }
  r->num_shared_first = r->shared_prefix.size() > 1 ? r->shared_prefix[1] : 0;
  halo_ptr.push_back((int32_t)slot);
  r->num_shared = (int64_t)halo_node.size();
  r->num_halo   = slot;
  r->num_nodes  = num_nodes;
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_tgt, n * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_halo_node, halo_node.size() * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&r->d_halo_ptr, halo_ptr.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, r->d_tgt, tgt.data(), n * sizeof(int32_t)));
  if (!halo_node.empty()) B200_CALL(b200_h2d(ceed, r->d_halo_node, halo_node.data(), halo_node.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, r->d_halo_ptr, halo_ptr.data(), halo_ptr.size() * sizeof(int32_t)));
  // dense form (see B200Restriction_::d_halo_cnt): only for an unpartitioned restriction -- the slots are then numbered in ascending
  // L-index order -- whose nodes have at most 255 halo slots each
  r->dense_n = 0;
  if (num_parts == 1 && r->num_shared > 0 && getenv("CEED_B200_DENSE_FINALIZE")) {  // (built only when the option is in play: measured slower)
    const int64_t dense_n = (int64_t)halo_node.back() + 1, padded = (dense_n + kDenseBlockNodes - 1) / kDenseBlockNodes * kDenseBlockNodes;
    std::vector<uint8_t> cnt(padded, 0);
    std::vector<int32_t> base(padded / kDenseBlockNodes, 0);
    bool                 ok = true;
    for (int64_t i = 0; i < r->num_shared && ok; i++) {
      const int32_t k = halo_ptr[i + 1] - halo_ptr[i];
      ok              = k <= 255 && (i == 0 || halo_node[i] > halo_node[i - 1]);
      cnt[halo_node[i]] = (uint8_t)k;
    }
    if (ok) {
      int64_t run = 0;
      for (int64_t b = 0; b < (int64_t)base.size(); b++) {
        base[b] = (int32_t)run;
        for (int64_t l = b * kDenseBlockNodes; l < (b + 1) * kDenseBlockNodes; l++) run += cnt[l];
      }
      B200_CALL(b200_dmalloc(ceed, (void **)&r->d_halo_cnt, cnt.size()));
      B200_CALL(b200_dmalloc(ceed, (void **)&r->d_halo_block_base, base.size() * sizeof(int32_t)));
      B200_CALL(b200_h2d(ceed, r->d_halo_cnt, cnt.data(), cnt.size()));
      B200_CALL(b200_h2d(ceed, r->d_halo_block_base, base.data(), base.size() * sizeof(int32_t)));
      r->dense_n = padded;
    }
  }
  r->owner_built = true;
  return B200_SUCCESS;
}

// Run scatter tables (see B200RunScatter): run of group w = elements [w * ne / G, (w + 1) * ne / G), walked in nb = ceil(len / E)
// iterations; element e of the run sits in iteration (e - s) % nb.  Host-side, O(E-entries).
int b200_restriction_build_runs(B200Restriction r, int num_groups, int group_elems, B200RunScatter *out) {
  B200Ceed ceed = r->ceed;
  if (out->built && out->num_groups == num_groups && out->group_elems == group_elems) return B200_SUCCESS;
  b200_run_scatter_free(ceed, out);
  B200_CALL(check_table_range(r));
  B200_CHECK(r->l_size < (int64_t)B200_RUN_RMW_BIT, ceed, B200_ERROR_UNSUPPORTED, "run scatter tables need L-vectors shorter than 2^30");
  B200_CHECK(num_groups > 0 && group_elems > 0, ceed, B200_ERROR_DIMENSION, "invalid run shape");
  HostTranspose t;
  build_host_transpose(r, t);
  const int64_t        ne = r->num_elem, es = r->elem_size, n = ne * es, num_nodes = (int64_t)t.lvec_indices.size();
  std::vector<int32_t> grp(ne), iter(ne);
  for (int64_t w = 0; w < num_groups; w++) {
    const int64_t s = w * ne / num_groups, s1 = (w + 1) * ne / num_groups, len = s1 - s;
    if (len <= 0) continue;
    const int64_t nb = (len + group_elems - 1) / group_elems;
    for (int64_t e = s; e < s1; e++) grp[e] = (int32_t)w, iter[e] = (int32_t)((e - s) % nb);
  }
  std::vector<int32_t> tgt(n), halo_node, halo_ptr;
  int64_t              slot = 0, num_rmw = 0;
  for (int64_t row = 0; row < num_nodes; row++) {
    const int32_t begin = t.t_offsets[row], end = t.t_offsets[row + 1];
    const int32_t l     = t.lvec_indices[row];
    tgt[t.t_indices[begin]] = l;  // owner: plain store
    bool    chain = true, any_halo = false;
    int64_t prev  = t.t_indices[begin] / es;
    const int32_t owner_group = grp[prev];
    for (int32_t j = begin + 1; j < end; j++) {
      const int64_t e = t.t_indices[j] / es;
      chain           = chain && grp[e] == owner_group && iter[e] > iter[prev];
      prev            = e;
      if (chain) {
        tgt[t.t_indices[j]] = l | B200_RUN_RMW_BIT;
        num_rmw++;
      } else {
        if (!any_halo) {
          halo_node.push_back(l);
          halo_ptr.push_back((int32_t)slot);
          any_halo = true;
        }
        tgt[t.t_indices[j]] = ~(int32_t)(slot++);
      }
    }
  }
  halo_ptr.push_back((int32_t)slot);
  out->num_shared = (int64_t)halo_node.size();
  out->num_halo   = slot;
  out->num_rmw    = num_rmw;
  B200_CALL(b200_dmalloc(ceed, (void **)&out->d_tgt, n * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&out->d_halo_node, (halo_node.size() + 1) * sizeof(int32_t)));
  B200_CALL(b200_dmalloc(ceed, (void **)&out->d_halo_ptr, halo_ptr.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, out->d_tgt, tgt.data(), n * sizeof(int32_t)));
  if (!halo_node.empty()) B200_CALL(b200_h2d(ceed, out->d_halo_node, halo_node.data(), halo_node.size() * sizeof(int32_t)));
  B200_CALL(b200_h2d(ceed, out->d_halo_ptr, halo_ptr.data(), halo_ptr.size() * sizeof(int32_t)));
  out->num_groups = num_groups, out->group_elems = group_elems;
  out->built = true;
  return B200_SUCCESS;
}

void b200_run_scatter_free(B200Ceed ceed, B200RunScatter *t) {
  b200_dfree(ceed, t->d_tgt);
  b200_dfree(ceed, t->d_halo_node);
  b200_dfree(ceed, t->d_halo_ptr);
  *t = B200RunScatter();
}

// Ordered scatter tables for element groups of `group_elems` consecutive elements (see B200OrderedScatter).
// Halo image of one shared node: [node id (as a 64-bit integer)][contribution of toucher 0]...[contribution of the last toucher].
// tgt of an E-entry: >= 0 plain L-index; < 0: x = ~tgt, bits 0..26 = slot of this entry's contribution, bit 27 = "last toucher:
// completes the node", bits 28..30 = number of touchers - 2 (valid on the last toucher).
namespace {
__global__ void k_ordered_write_ids(double *__restrict__ halo, const int32_t *__restrict__ id_slot, const int32_t *__restrict__ id_node, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<long long *>(halo)[id_slot[i]] = id_node[i];
}
}  // namespace

int b200_restriction_build_ordered(B200Restriction r, int group_elems, B200OrderedScatter *out) {
  B200Ceed      ceed = r->ceed;
  B200_CALL(check_table_range(r));
  HostTranspose t;
  build_host_transpose(r, t);
  const int64_t n          = (int64_t)r->num_elem * r->elem_size;
  const int64_t num_nodes  = (int64_t)t.lvec_indices.size();
  const int64_t num_groups = ((int64_t)r->num_elem + group_elems - 1) / group_elems;

  int64_t       num_halo = 0, num_shared = 0;
  int32_t       max_cnt = 0;
  for (int64_t row = 0; row < num_nodes; row++) {
    const int32_t cnt = t.t_offsets[row + 1] - t.t_offsets[row];
    if (cnt > 1) num_halo += cnt + 1, num_shared++, max_cnt = std::max(max_cnt, cnt);
  }
  out->supported = num_halo < (1LL << 27) && max_cnt <= 9;
  if (!out->supported) return B200_SUCCESS;  // the caller falls back to the two-pass deterministic scheme
  std::vector<int32_t>              tgt(n), id_slot, id_node;
  std::vector<std::vector<int32_t>> pred(num_groups);
  id_slot.reserve(num_shared);
  id_node.reserve(num_shared);
  int64_t       slot          = 0;
  const int64_t group_entries = (int64_t)group_elems * r->elem_size;
  for (int64_t row = 0; row < num_nodes; row++) {
    const int32_t begin = t.t_offsets[row], end = t.t_offsets[row + 1], cnt = end - begin;
    if (cnt == 1) {
      tgt[t.t_indices[begin]] = t.lvec_indices[row];
      continue;
    }
    id_slot.push_back((int32_t)slot++);
    id_node.push_back(t.lvec_indices[row]);
    const int32_t last       = t.t_indices[end - 1];
    const int32_t last_group = (int32_t)(last / group_entries);
    for (int32_t k = begin; k < end - 1; k++) {
      tgt[t.t_indices[k]] = ~(int32_t)(slot++);
      const int32_t g = (int32_t)(t.t_indices[k] / group_entries);
      if (g != last_group && (pred[last_group].empty() || pred[last_group].back() != g)) pred[last_group].push_back(g);
    }
    tgt[last] = ~(int32_t)((int32_t)(slot++) | (1 << 27) | ((cnt - 2) << 28));
  }
  std::vector<int32_t> pred_ptr(num_groups + 1, 0), pred_idx;
  for (int64_t g = 0; g < num_groups; g++) {
    std::sort(pred[g].begin(), pred[g].end());
    pred[g].erase(std::unique(pred[g].begin(), pred[g].end()), pred[g].end());
    pred_ptr[g + 1] = pred_ptr[g] + (int32_t)pred[g].size();
    pred_idx.insert(pred_idx.end(), pred[g].begin(), pred[g].end());
  }
  out->num_shared = num_shared, out->num_halo = num_halo, out->num_groups = num_groups, out->num_pred = (int64_t)pred_idx.size();
  auto upload = [&](int32_t **d, const std::vector<int32_t> &h) -> int {
    B200_CALL(b200_dmalloc(ceed, (void **)d, std::max<size_t>(h.size(), 1) * sizeof(int32_t)));
    if (!h.empty()) B200_CALL(b200_h2d(ceed, *d, h.data(), h.size() * sizeof(int32_t)));
    return B200_SUCCESS;
  };
  B200_CALL(upload(&out->d_tgt, tgt));
  B200_CALL(upload(&out->d_node, id_node));
  B200_CALL(upload(&out->d_ptr, id_slot));
  B200_CALL(upload(&out->d_pred_ptr, pred_ptr));
  B200_CALL(upload(&out->d_pred_idx, pred_idx));
  B200_CALL(upload(&out->d_flags, std::vector<int32_t>(num_groups, 0)));
  B200_CALL(upload(&out->d_sync, std::vector<int32_t>(2, 0)));
  r->num_nodes = num_nodes;
  return B200_SUCCESS;
}

// write the (static) node ids into a freshly allocated halo buffer
int b200_ordered_init_halo(B200Restriction r, const B200OrderedScatter *t, double *d_halo) {
  B200Ceed ceed = r->ceed;
  if (t->num_shared > 0 && !b200_compile_only()) LAUNCH(ceed, k_ordered_write_ids, t->num_shared, d_halo, t->d_ptr, t->d_node, t->num_shared);
  return B200_SUCCESS;
}

void b200_ordered_scatter_free(B200Ceed ceed, B200OrderedScatter *t) {
  for (int32_t **p : {&t->d_tgt, &t->d_node, &t->d_ptr, &t->d_pred_ptr, &t->d_pred_idx, &t->d_flags, &t->d_sync}) {
    b200_dfree(ceed, *p);
    *p = nullptr;
  }
}

// Diagnostics / host-logic tests: host copies of the scatter tables of an offset restriction.
extern "C" int ceedb200_restriction_debug_scatter_tables(B200Restriction r, int mode, int group_elems, int32_t *tgt, int64_t *counts, int32_t *pred_ptr,
                                                         int32_t *pred_idx, int64_t pred_capacity) {
  B200Ceed      ceed = r->ceed;
  const int64_t n    = (int64_t)r->num_elem * r->elem_size;
  B200_CHECK(!r->is_strided, ceed, B200_ERROR_UNSUPPORTED, "scatter tables exist for offset restrictions only");
  if (mode == B200_SCATTER_ORDERED) {
    B200OrderedScatter t;
    B200_CALL(b200_restriction_build_ordered(r, group_elems, &t));
    counts[0] = t.num_shared, counts[1] = t.num_halo, counts[2] = t.num_groups, counts[3] = t.num_pred, counts[4] = t.supported;
    if (t.supported) {
      B200_CALL(b200_d2h(ceed, tgt, t.d_tgt, n * sizeof(int32_t)));
      if (pred_ptr) B200_CALL(b200_d2h(ceed, pred_ptr, t.d_pred_ptr, (t.num_groups + 1) * sizeof(int32_t)));
      if (pred_idx && t.num_pred <= pred_capacity && t.num_pred > 0) B200_CALL(b200_d2h(ceed, pred_idx, t.d_pred_idx, t.num_pred * sizeof(int32_t)));
      b200_ordered_scatter_free(ceed, &t);
    }
  } else if (mode == 5) {  // owner/halo tables with K = group_elems equal element parts (streamed apply): pred_ptr receives shared_prefix, pred_idx halo_node
    std::vector<int32_t> ends;
    for (int c = 1; c <= group_elems; c++) ends.push_back((int32_t)((int64_t)r->num_elem * c / group_elems));
    B200_CALL(b200_restriction_set_parts(r, ends));
    B200_CALL(b200_restriction_build_owner(r));
    counts[0] = r->num_shared, counts[1] = r->num_halo, counts[2] = (int64_t)r->shared_prefix.size(), counts[3] = 0, counts[4] = 1;
    B200_CALL(b200_d2h(ceed, tgt, r->d_tgt, n * sizeof(int32_t)));
    for (size_t i = 0; i < r->shared_prefix.size() && pred_ptr; i++) pred_ptr[i] = (int32_t)r->shared_prefix[i];
    if (pred_idx && r->num_shared <= pred_capacity && r->num_shared > 0) B200_CALL(b200_d2h(ceed, pred_idx, r->d_halo_node, r->num_shared * sizeof(int32_t)));
  } else if (mode == 4) {  // run scatter: group_elems = elements per iteration, pred_capacity = number of groups (warps) of the launch
    B200RunScatter t;
    B200_CALL(b200_restriction_build_runs(r, (int)pred_capacity, group_elems, &t));
    counts[0] = t.num_shared, counts[1] = t.num_halo, counts[2] = t.num_groups, counts[3] = t.num_rmw, counts[4] = 1;
    B200_CALL(b200_d2h(ceed, tgt, t.d_tgt, n * sizeof(int32_t)));
    b200_run_scatter_free(ceed, &t);
  } else {
    B200_CALL(b200_restriction_build_owner(r));
    counts[0] = r->num_shared, counts[1] = r->num_halo, counts[2] = 0, counts[3] = 0, counts[4] = 1;
    B200_CALL(b200_d2h(ceed, tgt, r->d_tgt, n * sizeof(int32_t)));
  }
  return B200_SUCCESS;
}

int b200_restriction_e_size(B200Restriction r, int64_t *e_size) {
  *e_size = (int64_t)r->num_elem * r->elem_size * r->num_comp;
  return B200_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ apply
int b200_restriction_apply_raw(B200Restriction r, int t_mode, const double *d_u, double *d_v) {
  B200Ceed      ceed      = r->ceed;
  const int64_t e_entries = (int64_t)r->num_elem * r->elem_size;
  if (e_entries == 0) return B200_SUCCESS;
  if (r->is_strided) {
    if (t_mode == B200_NOTRANSPOSE)
      LAUNCH(ceed, k_strided_notranspose, e_entries, d_u, d_v, (int64_t)r->num_elem, r->elem_size, r->num_comp, r->strides[0], r->strides[1],
             r->strides[2]);
    else
      LAUNCH(ceed, k_strided_transpose, e_entries, d_u, d_v, (int64_t)r->num_elem, r->elem_size, r->num_comp, r->strides[0], r->strides[1],
             r->strides[2]);
  } else if (t_mode == B200_NOTRANSPOSE) {
    LAUNCH(ceed, k_offset_notranspose, e_entries, r->d_offsets, d_u, d_v, e_entries, r->num_comp, r->comp_stride);
  } else {
    B200_CALL(b200_restriction_build_transpose(r));
    LAUNCH(ceed, k_offset_transpose, r->num_nodes, r->d_lvec_indices, r->d_t_offsets, r->d_t_indices, d_u, d_v, r->num_nodes, e_entries, r->num_comp,
           r->comp_stride);
  }
  return B200_SUCCESS;
}

// CEED_B200_DENSE_FINALIZE=0|1: the dense form of the finalize pass for whole-mesh applies (read at every apply: the tuning scripts toggle it)
static bool b200_dense_finalize_enabled() {
  const char *e = getenv("CEED_B200_DENSE_FINALIZE");
  return e ? atoi(e) != 0 : false;
}

int b200_halo_finalize(B200Restriction r, const double *d_halo, double *d_v, int part) {
  B200Ceed      ceed  = r->ceed;
  int64_t first = 0, count = r->num_shared;
  if (part >= 1) {
    B200_CHECK(part < (int)r->shared_prefix.size(), ceed, B200_ERROR_DIMENSION, "element part %d of %d", part, (int)r->shared_prefix.size() - 1);
    first = r->shared_prefix[part - 1];
    count = r->shared_prefix[part] - first;
  }
  if (count <= 0) return B200_SUCCESS;
  if (part == 0 && r->dense_n > 0 && b200_dense_finalize_enabled()) {
    // whole mesh, unpartitioned restriction: dense tables (v must cover the padded index range only where counts are non-zero)
    B200_CHECK(!b200_compile_only(), ceed, B200_ERROR_BACKEND, "CEED_B200_COMPILE_ONLY is set; kernels cannot run");
    k_halo_finalize_dense<<<(unsigned)(r->dense_n / kDenseBlockNodes), kDenseThreads, 0, ceed->stream>>>(r->d_halo_cnt, r->d_halo_block_base, d_halo, d_v, r->num_halo,
                                                                                                   r->num_comp, r->comp_stride);
    ceed->launch_count++;
    B200_CUDA(ceed, cudaGetLastError());
    return B200_SUCCESS;
  }
  LAUNCH(ceed, k_halo_finalize, count, r->d_halo_node + first, r->d_halo_ptr + first, d_halo, d_v, count, r->num_halo, r->num_comp, r->comp_stride);
  return B200_SUCCESS;
}

int b200_halo_finalize_lists(B200Restriction r, const int32_t *d_halo_node, const int32_t *d_halo_ptr, int64_t num_shared, int64_t num_halo, const double *d_halo,
                             double *d_v) {
  B200Ceed ceed = r->ceed;
  if (num_shared <= 0) return B200_SUCCESS;
  LAUNCH(ceed, k_halo_finalize, num_shared, d_halo_node, d_halo_ptr, d_halo, d_v, num_shared, num_halo, r->num_comp, r->comp_stride);
  return B200_SUCCESS;
}

// Element split of a partitioned mesh: elements [0, split_elem) are the ones touching the rank interface (mesh.Partition orders
// them first).  Must be set before the restriction is first used by a fused operator (the owner/halo tables depend on it).
extern "C" int ceedb200_restriction_set_split(B200Restriction r, b200_int split_elem) {
  B200Ceed ceed = r->ceed;
  B200_CHECK(!r->is_strided, ceed, B200_ERROR_UNSUPPORTED, "element split applies to offset restrictions");
  B200_CHECK(split_elem >= 0 && split_elem <= r->num_elem, ceed, B200_ERROR_DIMENSION, "element split %d outside [0, %d]", split_elem, r->num_elem);
  B200_CALL(b200_restriction_set_parts(r, {(int32_t)split_elem, (int32_t)r->num_elem}));  // (either part may be empty)
  r->split_elem = split_elem;
  return B200_SUCCESS;
}

int b200_restriction_set_parts(B200Restriction r, const std::vector<int32_t> &part_ends) {
  B200Ceed ceed = r->ceed;
  if (r->part_ends == part_ends) return B200_SUCCESS;
  if (r->owner_built) {
    b200_dfree(ceed, r->d_tgt);
    b200_dfree(ceed, r->d_halo_node);
    b200_dfree(ceed, r->d_halo_ptr);
    b200_dfree(ceed, r->d_halo_cnt);
    b200_dfree(ceed, r->d_halo_block_base);
    r->d_tgt = r->d_halo_node = r->d_halo_ptr = nullptr;
    r->d_halo_cnt = nullptr, r->d_halo_block_base = nullptr, r->dense_n = 0;
    r->owner_built = false;
  }
  r->part_ends  = part_ends;
  r->split_elem = -1;
  return B200_SUCCESS;
}

// CeedElemRestrictionApply: NoTranspose overwrites the E-vector, Transpose ADDS into the L-vector
// (interface/ceed-elemrestriction.c CeedElemRestrictionApply doc; backends/ref/ceed-ref-restriction.c:220-242).
extern "C" int ceedb200_restriction_apply(B200Restriction r, int t_mode, B200Vector u, B200Vector v) {
  B200Ceed ceed = r->ceed;
  int64_t  e_size;
  B200_CALL(b200_restriction_e_size(r, &e_size));
  if (t_mode == B200_NOTRANSPOSE) {
    B200_CHECK(u->length >= r->l_size && v->length >= e_size, ceed, B200_ERROR_DIMENSION,
               "Input/output vectors too small for restriction: need L %lld, E %lld", (long long)r->l_size, (long long)e_size);
  } else {
    B200_CHECK(v->length >= r->l_size && u->length >= e_size, ceed, B200_ERROR_DIMENSION,
               "Input/output vectors too small for transpose restriction: need L %lld, E %lld", (long long)r->l_size, (long long)e_size);
  }
  const double *d_u;
  double       *d_v;
  B200_CALL(b200_vector_device_read(u, &d_u));
  // NoTranspose writes every E-entry; a longer destination keeps its remaining data
  B200_CALL(b200_vector_device_write(v, &d_v, t_mode == B200_NOTRANSPOSE && v->length == e_size));
  return b200_restriction_apply_raw(r, t_mode, d_u, d_v);
}

extern "C" int ceedb200_restriction_apply_ptr(B200Restriction r, int t_mode, const b200_scalar *d_u, b200_scalar *d_v) {
  return b200_restriction_apply_raw(r, t_mode, d_u, d_v);
}

extern "C" int ceedb200_restriction_get_offsets(B200Restriction r, int mem_type, const b200_int **offsets) {
  B200_CHECK(!r->is_strided, r->ceed, B200_ERROR_UNSUPPORTED, "strided restriction has no offsets");
  *offsets = mem_type == B200_MEM_HOST ? r->h_offsets : r->d_offsets;
  return B200_SUCCESS;
}

extern "C" int ceedb200_restriction_get_e_layout(B200Restriction r, b200_int layout[3]) {
  layout[0] = 1;
  layout[1] = r->elem_size * r->num_elem;
  layout[2] = r->elem_size;
  return B200_SUCCESS;
}

extern "C" int ceedb200_restriction_get_info(B200Restriction r, b200_int *num_elem, b200_int *elem_size, b200_int *num_comp, b200_size *l_size,
                                             b200_size *e_size) {
  if (num_elem) *num_elem = r->num_elem;
  if (elem_size) *elem_size = r->elem_size;
  if (num_comp) *num_comp = r->num_comp;
  if (l_size) *l_size = r->l_size;
  if (e_size) *e_size = (int64_t)r->num_elem * r->elem_size * r->num_comp;
  return B200_SUCCESS;
}
