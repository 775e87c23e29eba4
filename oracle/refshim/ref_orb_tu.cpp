// oracle/refshim/ref_orb_tu.cpp -- ORACLE test infrastructure.  Compiles the UNMODIFIED /root/reference/src/ORBextractor.cc
// (found through -I$(REF)/src).  Two things are arranged from OUTSIDE the source:
//  1. its dead computeDescriptors() call (ORBextractor.cc:1091) is enabled through the clock() hook at the end of
//     include/sdpl_cvshim.hpp; computeDescriptors() is file-static, so the hook's body has to live in this translation unit;
//  2. DistributeOctTree sorts vector<pair<int, ExtractorNode*>> (ORBextractor.cc:662-676): equal sizes are ordered by the
//     HEAP ADDRESS of the list node -- the reference's result depends on the allocator's state.  The list nodes
//     of std::list<ExtractorNode> are therefore given a monotonic arena (addresses grow with creation order, nothing is
//     reused), which makes the compiled reference deterministic and equal to oracle decision (i) "later-created node sorts
//     higher".  SDPL_REF_HEAP=1 in the environment switches back to plain operator new to measure how often that matters.
#define SDPL_REF_ENABLE_ORB_DESCRIPTORS 1
#include "sdpl_cvshim.hpp"
#include <list>
#include <sys/mman.h>

namespace SDPL_SLAM { class ExtractorNode; }
namespace sdpl_ref_arena {
static char* g_base = 0;
static size_t g_off = 0;
static const size_t kBytes = (size_t)1 << 30;
static int g_heap = -1;
inline bool use_heap() {
  if (g_heap < 0) { const char* e = getenv("SDPL_REF_HEAP"); g_heap = (e && e[0] == '1') ? 1 : 0; }
  return g_heap == 1;
}
inline void* bump(size_t bytes) {
  if (use_heap()) return ::operator new(bytes);
  if (!g_base) g_base = (char*)mmap(0, kBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  bytes = (bytes + 15) & ~(size_t)15;
  if (g_off + bytes > kBytes) { fprintf(stderr, "sdpl ref arena exhausted\n"); abort(); }
  void* p = g_base + g_off;
  g_off += bytes;
  return p;
}
inline void release(void* p) { if (use_heap()) ::operator delete(p); }
}  // namespace sdpl_ref_arena
extern "C" void ref_orb_arena_reset() { sdpl_ref_arena::g_off = 0; }

namespace std {
template <> class allocator<_List_node<SDPL_SLAM::ExtractorNode> > {
 public:
  typedef _List_node<SDPL_SLAM::ExtractorNode> value_type;
  typedef value_type* pointer;
  typedef const value_type* const_pointer;
  typedef value_type& reference;
  typedef const value_type& const_reference;
  typedef size_t size_type;
  typedef ptrdiff_t difference_type;
  typedef true_type propagate_on_container_move_assignment;
  typedef true_type is_always_equal;
  template <class U> struct rebind { typedef allocator<U> other; };
  allocator() noexcept {}
  allocator(const allocator&) noexcept {}
  template <class U> allocator(const allocator<U>&) noexcept {}
  pointer allocate(size_type n, const void* = 0);
  void deallocate(pointer p, size_type);
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
  template <class U> void destroy(U* p) { p->~U(); }
  size_type max_size() const noexcept { return size_t(-1) / 64; }
};
}  // namespace std

#include "ORBextractor.cc"

inline std::allocator<std::_List_node<SDPL_SLAM::ExtractorNode> >::pointer
std::allocator<std::_List_node<SDPL_SLAM::ExtractorNode> >::allocate(size_type n, const void*) {
  return (pointer)sdpl_ref_arena::bump(n * sizeof(value_type));
}
inline void std::allocator<std::_List_node<SDPL_SLAM::ExtractorNode> >::deallocate(pointer p, size_type) {
  sdpl_ref_arena::release(p);
}

namespace sdpl_ref_hook {
void at_clock(cv::Mat& working, std::vector<cv::KeyPoint>& kps, cv::Mat& d, std::vector<cv::Point>& pat) {
  SDPL_SLAM::computeDescriptors(working, kps, d, pat);
}
}  // namespace sdpl_ref_hook
