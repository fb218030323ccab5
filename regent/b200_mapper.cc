// b200_mapper.cc — placement policy for Regent-FFT on a B200 box.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD: it needs Legion's headers (legion.h, mappers/default_mapper.h),
// which are not available in the build image.  It is written against the public
// DefaultMapper policy hooks the reference's own mapper uses (test/test_mapper.cc:33-58) and is compiled the
// way the reference compiles its mapper (test/test_mapper.rg:23-59: $CXX -O2 -Wall -Werror -shared -fPIC with
// the Legion include path).
//
// What it changes relative to the reference's FFTTestMapper, which forces EVERY instance into zero-copy
// (pinned host) memory so that a CPU-captured pointer is valid on the GPU (test/test_mapper.cc:45-58) and thereby
// makes every FFT pass stream over the host link:
//   * data regions of tasks that run on a GPU processor go to that GPU's FRAMEBUFFER (HBM): libfft_b200's passes
//     then run at HBM speed (it would otherwise stage host instances through HBM itself);
//   * the plan region (a handful of iface.plan structs, src/fft.rg:48-65) stays in zero-copy memory, which
//     satisfies get_plan's SYSTEM/REGDMA/Z_COPY assertion (src/fft.rg:165-169) from both CPU and GPU tasks;
//   * the points of an index launch over slabs (execute_plan_task(r_part[i], s_part[i], p),
//     test/fft_test.rg:299-302) are spread over the node's GPUs in order, slab i -> GPU i mod #GPUs, which is the
//     placement libfft_b200's slab plans assume (rank r on GPU r, fft_b200.h "multi-GPU slab transforms").
#include "b200_mapper.h"

#include <vector>

#include "legion.h"
#include "mappers/default_mapper.h"

using namespace Legion;
using namespace Legion::Mapping;

namespace {

class FFTB200Mapper : public DefaultMapper {
public:
  FFTB200Mapper(MapperRuntime *rt, Machine machine, Processor local, const char *name)
    : DefaultMapper(rt, machine, local, name) {
    // the GPUs of this address space, in a fixed order
    Machine::ProcessorQuery gpus(machine);
    gpus.only_kind(Processor::TOC_PROC).same_address_space_as(local);
    for (Machine::ProcessorQuery::iterator it = gpus.begin(); it != gpus.end(); ++it) local_gpu_list.push_back(*it);
  }

  // Regions of at most this many elements are treated as plan regions (one iface.plan per node at most).
  static const size_t PLAN_REGION_MAX_VOLUME = 4096;

  virtual Memory default_policy_select_target_memory(MapperContext ctx, Processor target_proc,
                                                     const RegionRequirement &req,
                                                     MemoryConstraint mc = MemoryConstraint()) {
    if (target_proc.kind() == Processor::TOC_PROC) {
      const Domain dom = runtime->get_index_space_domain(ctx, req.region.get_index_space());
      const bool plan_like = dom.get_volume() <= PLAN_REGION_MAX_VOLUME;
      const Memory::Kind want = plan_like ? Memory::Z_COPY_MEM : Memory::GPU_FB_MEM;
      Machine::MemoryQuery q(machine);
      q.only_kind(want).best_affinity_to(target_proc);
      if (q.count() > 0) return q.first();
    } else if (target_proc.kind() == Processor::LOC_PROC) {
      // CPU tasks that touch a plan region (make_plan, destroy_plan are inlined into CPU parents,
      // src/fft.rg:261, 624) must see the same instance the GPU tasks use
      const Domain dom = runtime->get_index_space_domain(ctx, req.region.get_index_space());
      if (dom.get_volume() <= PLAN_REGION_MAX_VOLUME) {
        Machine::MemoryQuery q(machine);
        q.only_kind(Memory::Z_COPY_MEM).has_affinity_to(target_proc);
        if (q.count() > 0) return q.first();
      }
    }
    return DefaultMapper::default_policy_select_target_memory(ctx, target_proc, req, mc);
  }

  // slab i of an index launch -> GPU i mod #GPUs of this node
  virtual void slice_task(const MapperContext ctx, const Task &task, const SliceTaskInput &input,
                          SliceTaskOutput &output) {
    if (local_gpu_list.empty() || task.target_proc.kind() != Processor::TOC_PROC || input.domain.get_dim() != 1) {
      DefaultMapper::slice_task(ctx, task, input, output);
      return;
    }
    const Rect<1> rect = input.domain;
    for (PointInRectIterator<1> p(rect); p(); p++) {
      TaskSlice slice;
      slice.domain = Domain(Rect<1>(*p, *p));
      slice.proc = local_gpu_list[(size_t)((*p)[0] - rect.lo[0]) % local_gpu_list.size()];
      slice.recurse = false;
      slice.stealable = false;
      output.slices.push_back(slice);
    }
  }

private:
  std::vector<Processor> local_gpu_list;
};

void create_mappers(Machine machine, Runtime *runtime, const std::set<Processor> &local_procs) {
  for (std::set<Processor>::const_iterator it = local_procs.begin(); it != local_procs.end(); ++it)
    runtime->replace_default_mapper(new FFTB200Mapper(runtime->get_mapper_runtime(), machine, *it, "fft_b200_mapper"), *it);
}

}  // namespace

void register_mappers(void) { Runtime::add_registration_callback(create_mappers); }
