// Stand-in for CUSP v0.5.1 (not vendored by the reference, absent from this image): just enough for the two
// typedefs at the top of GPU/detail/format.h, so that the reference header itself can be compiled by
// tests/abi_cxx/layout_check.cpp.  Test infrastructure only.
#pragma once
namespace cusp {
struct host_memory {};
struct device_memory {};
template <class I, class V, class M> struct csr_matrix {};
template <class I, class V, class M> struct coo_matrix {};
}  // namespace cusp
