// Internal accessors shared between session.cu and host_io.cpp.
#pragma once
#include <cstdint>
#include <vector>
struct arap_ctx;
namespace arapgs {
int session_num_nodes(arap_ctx* c);
int session_k(arap_ctx* c);
const std::vector<std::vector<uint32_t>>& session_blocks(arap_ctx* c);
const std::vector<int>& session_block_types(arap_ctx* c);
}  // namespace arapgs
