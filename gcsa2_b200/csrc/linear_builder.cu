/*
  linear_builder.cu -- construction of the index of a LINEAR reference on the device (sm_100a).

  The fixture generator of BASELINE.json configs[1] and configs[3]: a 3 Gbp order-128 index cannot come out of the
  in-memory host builder (builder.cpp keeps every path with its label, ~210 bytes per base), and the reference's own
  constructor is disk-based (GCSA::GCSA(InputGraph&, ...), src/gcsa.cpp:447-724).  For a graph that is one path

        #  ->  s[0]  ->  s[1]  ->  ...  ->  s[L-1]  ->  $          ($ -> # is the technical edge, include/gcsa/dbg.h:70-72)

  the result of that construction has a closed form, which is what is computed here:

    * every kmer starts at one text position and has one successor, so prefix doubling with pruning
      (src/path_graph.cpp:892-1032) only decides how long the label of a position grows before it is unique; the
      final path nodes (PathGraph::merge, src/path_graph.cpp:1154-1226) are the DISTINCT PREFIXES OF LENGTH
      K = k << doubling_steps OF THE SUFFIXES of T = # s $ ('$' repeated past the end: kmers that reach the sink are
      padded with the endmarker, src/files.cpp:272-282), in lexicographic order of the comp values ($ A C G T N #);
      suffixes that agree on K characters share a node, whose values are their start positions;
    * the LCP array holds the common prefix lengths of neighbouring nodes (src/path_graph.cpp:1204);
    * node v has a predecessor character c iff one of its positions is preceded by c (gcsa.cpp:573-588); the edge
      (c, v) leaves the node of the preceding position, so the out-degree of a node is the number of distinct
      nodes its positions continue into;
    * samples, SadaSparse / SadaCount counters as in src/gcsa.cpp:590-658: a position belongs to exactly one node
      here, so there are no redundant pointers.

  Mechanism: suffix sorting by radix sort.  Round 0 sorts all suffixes by their first 21 characters (3 bits per
  character in a 63-bit key, cub::DeviceRadixSort); the suffixes that still share a key are compacted and refined
  10 characters per round with keys (group, next 10 characters) until the K-th character, so a random reference is
  finished after one full-size sort and a repetitive one pays only for what is actually tied.  Everything else is
  one pass per array (ballot-packed bit vectors, scans, scatters).  The arrays come back as the same gcsa_b200_built
  the host builder returns; the two are bit-identical wherever both can run (tests/test_linear_builder.py).

  Limits: L + 2 < 2^32 - 1 positions (32-bit suffix numbers), K <= 255.
*/
#include <cuda_runtime.h>
#include <cub/cub.cuh>

void enginePoolTrim();      // engine.cu: the engine's idle scratch memory back to the driver

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

namespace {

typedef uint64_t u64;
typedef unsigned long long ull;
typedef uint32_t u32;
typedef uint8_t  u8;

constexpr int LB_SIGMA = GCSA_B200_SIGMA;
constexpr int LB_FIRST_CHARS = 21;      // characters in the key of round 0 (63 bits)
constexpr int LB_ROUND_CHARS = 10;      // characters per refinement round (30 bits below a 32-bit group number)
constexpr int LB_THREADS = 256;

// The text and the node_type of its positions (include/gcsa/support.h:443-471: id << 11 | offset).
struct LinearText
{
  const u8* T;                          // n comp values: T[0] = '#' (6), T[n - 1] = '$' (0)
  u64 n, node_len, first_id, source_value, sink_value;
  u32 K;
};

__device__ __forceinline__ u64 lb_value(const LinearText& t, u64 p)
{
  if(p == 0) { return t.source_value; }
  if(p == t.n - 1) { return t.sink_value; }
  u64 s = p - 1;
  return ((t.first_id + s / t.node_len) << 11) | (s % t.node_len);
}

// The text is circular through the technical edge $ -> #.
__device__ __forceinline__ u64 lb_pred(const LinearText& t, u64 p) { return (p == 0 ? t.n - 1 : p - 1); }
__device__ __forceinline__ u64 lb_succ(const LinearText& t, u64 p) { return (p == t.n - 1 ? 0 : p + 1); }

// `count` characters of the suffix at pos, starting at offset off, 3 bits each, first character most significant;
// characters past the end of the text or past the K-th character of the suffix read as '$' (0).
__device__ __forceinline__ u64 lb_pack(const LinearText& t, u64 pos, u32 off, int count)
{
  u64 key = 0;
  for(int j = 0; j < count; j++)
  {
    u64 q = pos + off + j;
    u64 c = (off + j < t.K && q < t.n ? t.T[q] : 0);
    key = (key << 3) | c;
  }
  return key;
}

__global__ void lb_text_kernel(const u8* seq, u64 L, u8* T)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < L + 2; i += stride)
  {
    T[i] = (i == 0 ? (u8)6 : (i == L + 1 ? (u8)0 : seq[i - 1]));
  }
}

// seq values must be comps 1..5 (bases and N): anything else would collide with the source / sink markers
__global__ void lb_check_kernel(const u8* seq, u64 L, u32* bad)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += stride)
  {
    if(seq[i] < 1 || seq[i] > 5) { atomicAdd(bad, 1u); }
  }
}

__global__ void lb_key0_kernel(LinearText t, u64* keys, u32* sa)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride)
  {
    keys[i] = lb_pack(t, i, 0, LB_FIRST_CHARS); sa[i] = (u32)i;
  }
}

// head[i] = 1 iff the sorted key at i differs from its left neighbour's
__global__ void lb_heads_kernel(const u64* keys, u64 m, u8* head)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    head[i] = (i == 0 || keys[i] != keys[i - 1] ? 1 : 0);
  }
}

// tied[i] = 1 iff element i is in a group of more than one element
__global__ void lb_tied_kernel(const u8* head, u64 m, u32* tied)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    tied[i] = (!head[i] || (i + 1 < m && !head[i + 1]) ? 1u : 0u);
  }
}

// Compacts the tied elements: where[i] = exclusive scan of tied[].  slot_in == nullptr: the slot is i itself.
__global__ void lb_compact_kernel(const u32* tied_scan, const u8* head, const u32* slot_in, const u32* pos_in, u64 m, u64 total,
                                  u32* slot_out, u32* pos_out, u32* head_out)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    u32 w = tied_scan[i];
    bool is_tied = (i + 1 < m ? tied_scan[i + 1] != w : total != w);
    if(is_tied)
    {
      slot_out[w] = (slot_in != nullptr ? slot_in[i] : (u32)i);
      pos_out[w] = pos_in[i];
      head_out[w] = head[i];
    }
  }
}

// grp_scan = inclusive scan of the head flags of the compacted elements
__global__ void lb_round_key_kernel(LinearText t, const u32* grp_scan, const u32* pos, u64 m, u32 off, u64* keys)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    keys[i] = ((u64)(grp_scan[i] - 1) << 32) | lb_pack(t, pos[i], off, LB_ROUND_CHARS);
  }
}

// After a round: the t-th sorted element goes to the t-th tied slot (groups are contiguous and keep their sizes).
__global__ void lb_writeback_kernel(const u32* slot, const u32* pos_sorted, const u8* head_new, u64 m, u32* sa, u8* head)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    sa[slot[i]] = pos_sorted[i];
    if(head_new[i]) { head[slot[i]] = 1; }
  }
}

__global__ void lb_widen_kernel(const u8* in, u64 m, u32* out)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) { out[i] = in[i]; }
}

// node_scan = inclusive scan of head[]: the node of slot i is node_scan[i] - 1
__global__ void lb_nodes_kernel(const u8* head, const u32* node_scan, const u32* sa, u64 n, u64 N, u32* node_first, u32* node_of_pos)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    u32 v = node_scan[i] - 1;
    if(head[i]) { node_first[v] = (u32)i; }
    node_of_pos[sa[i]] = v;
    if(i == 0) { node_first[N] = (u32)n; }
  }
}

// ORs bits into byte idx of an array of bytes whose other bytes may be written concurrently by the same means.
__device__ __forceinline__ void lb_or_byte(u8* base, u64 idx, u32 bits)
{
  u32* word = (u32*)(base + (idx & ~(u64)3));
  atomicOr(word, bits << (8 * (u32)(idx & 3)));
}

// Per slot: the predecessor character of the position goes into the node's mask; a position that forces its node
// to be sampled (src/gcsa.cpp:631-646: value at a multiple of the sample period, or not the successor of the
// value before it) sets bit 7.
__global__ void lb_slot_kernel(LinearText t, const u32* sa, const u32* node_scan, u64 sample_period, u8* npreds)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride)
  {
    u64 p = sa[i];
    u32 v = node_scan[i] - 1;
    u64 q = lb_pred(t, p);
    u32 bits = 1u << t.T[q];
    u64 value = lb_value(t, p);
    if(value % sample_period == 0 || lb_value(t, q) + 1 != value) { bits |= 0x80u; }
    lb_or_byte(npreds, v, bits);
  }
}

__global__ void lb_lcp_kernel(LinearText t, const u32* sa, const u32* node_first, u64 N, u8* lcp)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += stride)
  {
    u32 len = 0;
    if(v > 0)
    {
      u64 a = sa[node_first[v - 1]], b = sa[node_first[v]];
      while(len < t.K)
      {
        u8 x = (a + len < t.n ? t.T[a + len] : 0), y = (b + len < t.n ? t.T[b + len] : 0);
        if(x != y) { break; }
        len++;
      }
    }
    lcp[v] = (u8)len;
  }
}

// One bit vector from a per-node predicate, 32 nodes per ballot; counts the ones.
// which: 0..6 = predecessor character c, 7 = sampled[], 8 = more than one value
__global__ void lb_bits_kernel(const u8* npreds, const u8* sampled, const u32* node_first, u64 N, int which, u32* words, ull* ones)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  u64 rounds = (N + stride - 1) / stride;
  for(u64 r = 0; r < rounds; r++)
  {
    u64 v = r * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool bit = false;
    if(v < N)
    {
      if(which < LB_SIGMA) { bit = (npreds[v] >> which) & 1; }
      else if(which == 7) { bit = (sampled[v] != 0); }
      else { bit = (node_first[v + 1] - node_first[v] > 1); }
    }
    u32 word = __ballot_sync(0xFFFFFFFFu, bit);
    if((threadIdx.x & 31) == 0 && v < N)
    {
      words[v >> 5] = word;
      if(word != 0 && ones != nullptr) { atomicAdd(ones, (ull)__popc(word)); }
    }
  }
}

// Targets of the positions of the multi-valued nodes: (node << 32 | node of the next position)
__global__ void lb_pairs_kernel(LinearText t, const u32* slot, const u32* sa, const u32* node_scan, const u32* node_of_pos, u64 m, u64* pairs)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    u32 s = slot[i];
    pairs[i] = ((u64)(node_scan[s] - 1) << 32) | node_of_pos[lb_succ(t, sa[s])];
  }
}

__global__ void lb_outdeg_init_kernel(const u32* node_first, u64 N, u32* outdeg)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += stride)
  {
    outdeg[v] = (node_first[v + 1] - node_first[v] > 1 ? 0u : 1u);
  }
}

__global__ void lb_outdeg_pairs_kernel(const u64* pairs, u64 m, u32* outdeg)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
  {
    if(i == 0 || pairs[i] != pairs[i - 1]) { atomicAdd(outdeg + (pairs[i] >> 32), 1u); }
  }
}

__global__ void lb_widen64_kernel(const u32* in, u64 m, u64* out)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) { out[i] = in[i]; }
}

// sets bit (end[v] - 1) for every v with flag (flags == nullptr: all); end = inclusive scan
__global__ void lb_mark_ends_kernel(const u64* end, const u32* node_first, int multi_only, u64 N, ull* words)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += stride)
  {
    if(multi_only && node_first[v + 1] - node_first[v] <= 1) { continue; }
    u64 e = end[v];
    if(e == 0) { continue; }
    e--;
    atomicOr(words + (e >> 6), 1ull << (e & 63));
  }
}

// src/gcsa.cpp:621-646: a node is sampled if it has several predecessors, follows the endmarker, holds a value that
// must be sampled, or its values are not exactly the successors of its predecessor's values.
__global__ void lb_sampled_kernel(LinearText t, const u8* npreds, const u32* sa, const u32* node_first, const u32* node_of_pos, u64 N,
                                  u8* sampled, u64* sample_values, u64* extra)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += stride)
  {
    u32 mask = npreds[v];
    u32 nvals = node_first[v + 1] - node_first[v];
    bool s = (__popc(mask & 0x7Fu) > 1) || (mask & 1u) || (mask & 0x80u);
    if(!s)
    {
      u32 u = node_of_pos[lb_pred(t, sa[node_first[v]])];
      s = (node_first[u + 1] - node_first[u] != nvals);
    }
    sampled[v] = (s ? 1 : 0);
    sample_values[v] = (s ? nvals : 0);
    extra[v] = (nvals > 1 ? nvals - 1 : 0);
  }
}

// sample_end = inclusive scan of sample_values
__global__ void lb_samples_kernel(LinearText t, const u8* sampled, const u32* sa, const u32* node_scan, const u32* node_first,
                                  const u64* sample_end, u64* stored)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += stride)
  {
    u32 v = node_scan[i] - 1;
    if(!sampled[v]) { continue; }
    u32 nvals = node_first[v + 1] - node_first[v];
    stored[sample_end[v] - nvals + (i - node_first[v])] = lb_value(t, sa[i]);
  }
}

__global__ void lb_mark_sample_ends_kernel(const u8* sampled, const u64* sample_end, u64 N, ull* words)
{
  u64 stride = (u64)gridDim.x * blockDim.x;
  for(u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += stride)
  {
    if(!sampled[v]) { continue; }
    u64 e = sample_end[v] - 1;
    atomicOr(words + (e >> 6), 1ull << (e & 63));
  }
}

//------------------------------------------------------------------------------
// Host side
//------------------------------------------------------------------------------

inline int lb_grid(u64 n, int sm_count)
{
  u64 blocks = (n + LB_THREADS - 1) / LB_THREADS;
  return (int)std::max<u64>(1, std::min<u64>(blocks, (u64)sm_count * 8));
}

inline size_t lb_words(u64 bits) { return (size_t)((bits + 63) / 64) + 1; }

struct Arena
{
  std::vector<void*> live;
  std::string error;
  template<class T> T* get(size_t count)
  {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(count * sizeof(T), 256);
    cudaError_t e = cudaMalloc(&p, bytes);
    if(e != cudaSuccess) { error = std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes: " + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
    live.push_back(p);
    return (T*)p;
  }
  void drop(void* p)
  {
    if(p == nullptr) { return; }
    for(size_t i = 0; i < live.size(); i++) { if(live[i] == p) { live.erase(live.begin() + i); break; } }
    cudaFree(p);
  }
  ~Arena() { for(void* p : live) { cudaFree(p); } }
};

int lb_fail(int code, const std::string& msg) { gcsa_b200_internal_set_error(msg.c_str()); return code; }

#define LB_CUDA(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess) { \
  return lb_fail(GCSA_B200_ERR_CUDA, std::string("build_linear: " #expr ": ") + cudaGetErrorString(e_)); } } while(0)
#define LB_ALLOC(var, type, count) type* var = arena.get<type>(count); \
  if(var == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: " + arena.error); }

struct Timer
{
  bool on = (std::getenv("GCSA_B200_VERBOSE") != nullptr);
  void lap(const char* what)
  {
    if(!on) { return; }
    cudaDeviceSynchronize();
    static thread_local double t0 = 0;
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    double t = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if(what != nullptr) { std::fprintf(stderr, "[build_linear] %-28s %8.3f s\n", what, t - t0); }
    t0 = t;
  }
};

template<class T> int lb_download(const T* dev, size_t count, size_t alloc_count, T** out)
{
  T* p = (T*)std::calloc(std::max<size_t>(alloc_count, 1), sizeof(T));
  if(p == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: host allocation failed"); }
  *out = p;
  if(count > 0) { LB_CUDA(cudaMemcpy(p, dev, count * sizeof(T), cudaMemcpyDeviceToHost)); }
  return 0;
}

int buildLinear(const u8* seq, u64 L, int seq_on_device, u64 node_len, int k, int steps, u64 sample_period, int device, gcsa_b200_built* result)
{
  if(seq == nullptr || L == 0 || k < 1 || k > 16 || steps < 0 || steps > 4) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: bad argument"); }
  const u32 K = (u32)k << steps;
  if(K > 255) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: order above 255 (LCP values are bytes, include/gcsa/support.h:44)"); }
  if(node_len == 0 || node_len > 1024) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: node length must be in 1..1024 (node offsets have 10 bits)"); }
  const u64 n = L + 2;
  if(n >= 0xFFFFFFFFull) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: more than 2^32 - 3 bases"); }
  if(sample_period == 0) { sample_period = 64; }

  int n_dev = 0;
  if(cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) { cudaGetLastError(); return lb_fail(GCSA_B200_ERR_CUDA, "build_linear: no CUDA device available"); }
  if(device < 0 || device >= n_dev) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: bad device ordinal"); }
  int prev_device = 0;
  LB_CUDA(cudaGetDevice(&prev_device));
  LB_CUDA(cudaSetDevice(device));
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore = { prev_device };
  enginePoolTrim();           // construction needs the device's memory: scratch kept from earlier query batches goes back first
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);

  Arena arena;
  Timer timer; timer.lap(nullptr);
  #define GRID(count) lb_grid((count), sm_count), LB_THREADS

  // ---- the text ----
  LB_ALLOC(T, u8, n + 8);
  {
    const u8* d_seq = seq;
    u8* staged = nullptr;
    if(!seq_on_device)
    {
      staged = arena.get<u8>(L);
      if(staged == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: " + arena.error); }
      LB_CUDA(cudaMemcpy(staged, seq, L, cudaMemcpyHostToDevice));
      d_seq = staged;
    }
    LB_ALLOC(d_bad, u32, 1);
    LB_CUDA(cudaMemset(d_bad, 0, sizeof(u32)));
    lb_check_kernel<<<GRID(L)>>>(d_seq, L, d_bad);
    lb_text_kernel<<<GRID(n)>>>(d_seq, L, T);
    u32 bad = 0;
    LB_CUDA(cudaMemcpy(&bad, d_bad, sizeof(u32), cudaMemcpyDeviceToHost));
    arena.drop(d_bad); arena.drop(staged);
    if(bad != 0) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: the sequence must consist of comp values 1..5 (A C G T N)"); }
  }
  LinearText t;
  t.T = T; t.n = n; t.node_len = node_len; t.first_id = 2; t.K = K;
  t.source_value = (u64)1 << 11;
  t.sink_value = (t.first_id + (L + node_len - 1) / node_len) << 11;

  // ---- round 0: all suffixes by their first 21 characters ----
  LB_ALLOC(sa, u32, n);
  LB_ALLOC(head, u8, n);
  {
    LB_ALLOC(keys_in, u64, n);
    LB_ALLOC(keys_out, u64, n);
    LB_ALLOC(sa_in, u32, n);
    lb_key0_kernel<<<GRID(n)>>>(t, keys_in, sa_in);
    size_t tmp_bytes = 0;
    LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, sa_in, sa, n, 0, 3 * LB_FIRST_CHARS));
    LB_ALLOC(tmp, u8, tmp_bytes);
    LB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, sa_in, sa, n, 0, 3 * LB_FIRST_CHARS));
    lb_heads_kernel<<<GRID(n)>>>(keys_out, n, head);
    LB_CUDA(cudaDeviceSynchronize());
    arena.drop(tmp); arena.drop(keys_in); arena.drop(keys_out); arena.drop(sa_in);
  }
  timer.lap("round 0 (21 characters)");

  // ---- refinement: the suffixes that are still tied, 10 characters per round ----
  u32* tied_slot = nullptr;             // slots of the positions in multi-valued nodes after the last round
  u64 tied = 0;
  {
    u32 *slot = nullptr, *pos = nullptr, *hd = nullptr;       // the tied elements: slot in sa[], suffix, head flag (u32 for the scan)
    u64 m = 0;
    {
      LB_ALLOC(flags, u32, n + 1);
      lb_tied_kernel<<<GRID(n)>>>(head, n, flags);
      size_t tmp_bytes = 0;
      LB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, flags, n + 1));
      LB_ALLOC(tmp, u8, tmp_bytes);
      LB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, flags, n + 1));
      u32 total = 0;
      LB_CUDA(cudaMemcpy(&total, flags + n, sizeof(u32), cudaMemcpyDeviceToHost));
      m = total;
      slot = arena.get<u32>(m); pos = arena.get<u32>(m); hd = arena.get<u32>(m);
      if(slot == nullptr || pos == nullptr || hd == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: " + arena.error); }
      if(m > 0) { lb_compact_kernel<<<GRID(n)>>>(flags, head, nullptr, sa, n, m, slot, pos, hd); }
      LB_CUDA(cudaDeviceSynchronize());
      arena.drop(tmp); arena.drop(flags);
    }
    for(u32 off = LB_FIRST_CHARS; off < K && m > 0; off += LB_ROUND_CHARS)
    {
      size_t tmp_bytes = 0, scan_bytes = 0;
      LB_ALLOC(keys_in, u64, m);
      LB_ALLOC(keys_out, u64, m);
      LB_ALLOC(pos_out, u32, m);
      LB_ALLOC(head_new, u8, m);
      LB_ALLOC(flags, u32, m + 1);
      LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, hd, hd, m));
      LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, pos, pos_out, m, 0, 64));
      tmp_bytes = std::max(tmp_bytes, scan_bytes);
      LB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, flags, m + 1));
      tmp_bytes = std::max(tmp_bytes, scan_bytes);
      LB_ALLOC(tmp, u8, tmp_bytes);
      LB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, hd, hd, m));                      // group numbers (+1)
      lb_round_key_kernel<<<GRID(m)>>>(t, hd, pos, m, off, keys_in);
      LB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, pos, pos_out, m, 0, 64));
      lb_heads_kernel<<<GRID(m)>>>(keys_out, m, head_new);
      lb_writeback_kernel<<<GRID(m)>>>(slot, pos_out, head_new, m, sa, head);
      lb_tied_kernel<<<GRID(m)>>>(head_new, m, flags);
      LB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, flags, m + 1));
      u32 total = 0;
      LB_CUDA(cudaMemcpy(&total, flags + m, sizeof(u32), cudaMemcpyDeviceToHost));
      u32 *slot2 = arena.get<u32>(total), *pos2 = arena.get<u32>(total), *hd2 = arena.get<u32>(total);
      if(slot2 == nullptr || pos2 == nullptr || hd2 == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: " + arena.error); }
      if(total > 0) { lb_compact_kernel<<<GRID(m)>>>(flags, head_new, slot, pos_out, m, total, slot2, pos2, hd2); }
      LB_CUDA(cudaDeviceSynchronize());
      arena.drop(tmp); arena.drop(keys_in); arena.drop(keys_out); arena.drop(pos_out); arena.drop(head_new); arena.drop(flags);
      arena.drop(slot); arena.drop(pos); arena.drop(hd);
      slot = slot2; pos = pos2; hd = hd2; m = total;
    }
    arena.drop(pos); arena.drop(hd);
    tied_slot = slot; tied = m;
  }
  timer.lap("refinement rounds");

  // ---- nodes ----
  LB_ALLOC(node_scan, u32, n);
  u64 N = 0;
  {
    lb_widen_kernel<<<GRID(n)>>>(head, n, node_scan);
    size_t tmp_bytes = 0;
    LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, node_scan, node_scan, n));
    LB_ALLOC(tmp, u8, tmp_bytes);
    LB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, node_scan, node_scan, n));
    u32 last = 0;
    LB_CUDA(cudaMemcpy(&last, node_scan + (n - 1), sizeof(u32), cudaMemcpyDeviceToHost));
    N = last;
    arena.drop(tmp);
  }
  LB_ALLOC(node_first, u32, N + 1);
  LB_ALLOC(node_of_pos, u32, n);
  lb_nodes_kernel<<<GRID(n)>>>(head, node_scan, sa, n, N, node_first, node_of_pos);
  arena.drop(head); head = nullptr;

  LB_ALLOC(npreds, u8, N + 4);
  LB_CUDA(cudaMemset(npreds, 0, N + 4));
  lb_slot_kernel<<<GRID(n)>>>(t, sa, node_scan, sample_period, npreds);
  LB_ALLOC(lcp, u8, N);
  lb_lcp_kernel<<<GRID(N)>>>(t, sa, node_first, N, lcp);
  timer.lap("nodes, masks, lcp");

  gcsa_flat_index& f = result->index;
  f.path_nodes = N; f.order = K; f.sigma = LB_SIGMA; f.fast_chars = GCSA_B200_FAST_CHARS;
  gcsa_b200_default_char2comp(f.char2comp);
  int rc = 0;
  #define LB_RC(expr) do { rc = (expr); if(rc) { return rc; } } while(0)
  LB_RC(lb_download(lcp, N, N + 1, &result->lcp));
  result->lcp_size = N;
  arena.drop(lcp);

  // ---- BWT bit vectors and C ----
  const size_t node_words = lb_words(N);
  {
    LB_ALLOC(words, u32, 2 * node_words);
    LB_ALLOC(d_ones, ull, LB_SIGMA);
    LB_CUDA(cudaMemset(d_ones, 0, LB_SIGMA * sizeof(ull)));
    for(int c = 0; c < LB_SIGMA; c++)
    {
      LB_CUDA(cudaMemset(words, 0, 2 * node_words * sizeof(u32)));
      lb_bits_kernel<<<GRID(N)>>>(npreds, nullptr, node_first, N, c, words, d_ones + c);
      u64* host = nullptr;
      LB_RC(lb_download((const u64*)words, node_words, node_words, &host));
      f.bwt[c] = host;
    }
    ull ones[LB_SIGMA];
    LB_CUDA(cudaMemcpy(ones, d_ones, sizeof(ones), cudaMemcpyDeviceToHost));
    f.C[0] = 0;
    for(int c = 0; c < LB_SIGMA; c++) { f.C[c + 1] = f.C[c] + ones[c]; }
    arena.drop(words); arena.drop(d_ones);
  }
  timer.lap("bwt");

  // ---- edges: out-degrees, then the unary code of them ----
  {
    LB_ALLOC(outdeg, u32, N);
    lb_outdeg_init_kernel<<<GRID(N)>>>(node_first, N, outdeg);
    if(tied > 0)
    {
      LB_ALLOC(pairs_in, u64, tied);
      LB_ALLOC(pairs_out, u64, tied);
      lb_pairs_kernel<<<GRID(tied)>>>(t, tied_slot, sa, node_scan, node_of_pos, tied, pairs_in);
      size_t tmp_bytes = 0;
      LB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, pairs_in, pairs_out, tied, 0, 64));
      LB_ALLOC(tmp, u8, tmp_bytes);
      LB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, pairs_in, pairs_out, tied, 0, 64));
      lb_outdeg_pairs_kernel<<<GRID(tied)>>>(pairs_out, tied, outdeg);
      LB_CUDA(cudaDeviceSynchronize());
      arena.drop(tmp); arena.drop(pairs_in); arena.drop(pairs_out);
    }
    LB_ALLOC(end, u64, N);
    lb_widen64_kernel<<<GRID(N)>>>(outdeg, N, end);
    arena.drop(outdeg);
    size_t tmp_bytes = 0;
    LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, end, end, N));
    LB_ALLOC(tmp, u8, tmp_bytes);
    LB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, end, end, N));
    u64 edge_count = 0;
    LB_CUDA(cudaMemcpy(&edge_count, end + (N - 1), sizeof(u64), cudaMemcpyDeviceToHost));
    f.edge_count = edge_count;
    const size_t edge_words = lb_words(edge_count);
    LB_ALLOC(words, ull, edge_words);
    LB_CUDA(cudaMemset(words, 0, edge_words * sizeof(ull)));
    lb_mark_ends_kernel<<<GRID(N)>>>(end, node_first, 0, N, words);
    u64* host = nullptr;
    LB_RC(lb_download((const u64*)words, edge_words, edge_words, &host));
    f.edges = host;
    arena.drop(tmp); arena.drop(end); arena.drop(words);
  }
  arena.drop(tied_slot);
  timer.lap("edges");

  // ---- samples and the counting structures ----
  {
    LB_ALLOC(sampled, u8, N);
    LB_ALLOC(sample_end, u64, N);
    LB_ALLOC(extra_end, u64, N);
    lb_sampled_kernel<<<GRID(N)>>>(t, npreds, sa, node_first, node_of_pos, N, sampled, sample_end, extra_end);
    size_t tmp_bytes = 0;
    LB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, sample_end, sample_end, N));
    LB_ALLOC(tmp, u8, tmp_bytes);
    LB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, sample_end, sample_end, N));
    LB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, extra_end, extra_end, N));
    u64 sample_count = 0, extra_len = 0;
    LB_CUDA(cudaMemcpy(&sample_count, sample_end + (N - 1), sizeof(u64), cudaMemcpyDeviceToHost));
    LB_CUDA(cudaMemcpy(&extra_len, extra_end + (N - 1), sizeof(u64), cudaMemcpyDeviceToHost));
    f.sample_count = sample_count; f.extra_values_len = extra_len;

    LB_ALLOC(words, u32, 2 * node_words);
    u64* host = nullptr;
    LB_CUDA(cudaMemset(words, 0, 2 * node_words * sizeof(u32)));
    lb_bits_kernel<<<GRID(N)>>>(npreds, sampled, node_first, N, 7, words, nullptr);
    LB_RC(lb_download((const u64*)words, node_words, node_words, &host)); f.sampled_paths = host;
    LB_CUDA(cudaMemset(words, 0, 2 * node_words * sizeof(u32)));
    lb_bits_kernel<<<GRID(N)>>>(npreds, sampled, node_first, N, 8, words, nullptr);
    LB_RC(lb_download((const u64*)words, node_words, node_words, &host)); f.extra_filter = host;
    arena.drop(words);

    LB_ALLOC(stored, u64, sample_count);
    lb_samples_kernel<<<GRID(n)>>>(t, sampled, sa, node_scan, node_first, sample_end, stored);
    LB_RC(lb_download((const u64*)stored, sample_count, sample_count + 1, &host)); f.stored_samples = host;
    arena.drop(stored);

    const size_t sample_words = lb_words(sample_count), extra_words = lb_words(extra_len);
    LB_ALLOC(swords, ull, sample_words);
    LB_CUDA(cudaMemset(swords, 0, sample_words * sizeof(ull)));
    lb_mark_sample_ends_kernel<<<GRID(N)>>>(sampled, sample_end, N, swords);
    LB_RC(lb_download((const u64*)swords, sample_words, sample_words, &host)); f.samples = host;
    LB_ALLOC(xwords, ull, extra_words);
    LB_CUDA(cudaMemset(xwords, 0, extra_words * sizeof(ull)));
    lb_mark_ends_kernel<<<GRID(N)>>>(extra_end, node_first, 1, N, xwords);
    LB_RC(lb_download((const u64*)xwords, extra_words, extra_words, &host)); f.extra_values = host;
    LB_CUDA(cudaDeviceSynchronize());
  }
  // SadaCount (support.h:264-279): no position occurs in two nodes, so every counter is 0 -> N - 1 ones
  {
    u64 len = (N > 0 ? N - 1 : 0);
    size_t words = lb_words(len);
    u64* host = (u64*)std::calloc(words, sizeof(u64));
    if(host == nullptr) { return lb_fail(GCSA_B200_ERR_NOMEM, "build_linear: host allocation failed"); }
    for(u64 w = 0; w < len / 64; w++) { host[w] = ~0ull; }
    if(len % 64) { host[len / 64] = (1ull << (len % 64)) - 1; }
    f.redundant = host; f.redundant_len = len;
  }
  timer.lap("samples, counters");
  #undef GRID
  #undef LB_RC

  // every edge has a source and a target: the two counts must agree (GCSAHeader::edges)
  if(f.edge_count != f.C[LB_SIGMA]) { return lb_fail(GCSA_B200_ERR_INCONSISTENT, "build_linear: out-degrees and predecessor sets disagree"); }
  return 0;
}

} // namespace

extern "C" {

int gcsa_b200_build_linear(const uint8_t* sequence, uint64_t length, int sequence_on_device, uint64_t node_length,
                           int kmer_length, int doubling_steps, uint64_t sample_period, int device, gcsa_b200_built* result)
{
  if(result == nullptr) { return lb_fail(GCSA_B200_ERR_INVALID, "build_linear: null result"); }
  std::memset(result, 0, sizeof(*result));
  int rc = buildLinear(sequence, length, sequence_on_device, node_length, kmer_length, doubling_steps, sample_period, device, result);
  if(rc != 0) { gcsa_b200_built_free(result); }
  return rc;
}

} // extern "C"
