/*
 * dcsim_advance_g8.cu — the event loop with 8 lanes per replica: a warp carries 4 replicas on aligned lane groups.
 *
 * Why: most of an event's instructions are warp-wide regardless of how many lanes do useful work — the pop-min, the
 * per-DC sweep (4 to 8 useful lanes), the handler on the group's lane 0.  With 4 replicas per warp the warp-wide part is
 * shared by 4 events, and handlers of the same kind run side by side (different kinds serialise).  Everything else is
 * the same source: dcsim_core.cuh is written against DCSIM_LANES, and every collective is group-relative.
 */
#define DCSIM_LANES 8
#define DCSIM_ADV_SUFFIX _g8
#include "dcsim_advance_impl.cuh"
