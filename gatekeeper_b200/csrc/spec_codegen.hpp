// The constraint netlist as generated CUDA C++ (the "specialised" evaluation kernel).
//
// gk_eval_kernel (tile_kernel.cuh) INTERPRETS the netlist: every node is a bit column of a 512-object tile in shared memory, every
// op a work item some warp decodes, ~400 warp-instructions per object at the benchmark's 50 constraints.  The netlist is known when
// the constraint set is compiled, so the engine also writes it out as straight-line code -- one thread per object, every node a
// 32-bit mask register (bit j = row j of THIS object in the node's scope), constants as immediates, EXISTS a `!= 0`, the
// object-major bitmap words assembled in registers without a transpose -- and NVRTC compiles that text for sm_100a when the first
// large batch is evaluated (kernels.cu).  Objects with more than 32 rows in some scope do not fit the mask registers: their tiles
// are handed to the interpreter (GkKParams::tile_list).
#pragma once
#include <string>

#include "engine.hpp"

namespace gk {

struct SpecSource {
  std::string src;        // one self-contained translation unit: program.h + vm_core.h + gk_spec_object() + gk_spec_kernel
  uint32_t words = 1;     // bitmap words per object
  size_t smem = 0;        // dynamic shared memory of a launch (column / scope tables, enforcement-point mask, totals)
  size_t n_fast = 0, n_generic = 0;   // atoms emitted as immediates / as calls of the generic gk_atom()
};

SpecSource spec_codegen(const Compiled& c);

}  // namespace gk
