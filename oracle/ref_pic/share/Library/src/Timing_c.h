// Stand-in for SWMF share/Library/src/Timing_c.h (un-vendored), OURS: the timing hooks are no-ops.
#pragma once
inline void timing_start(const char *) {}
inline void timing_stop(const char *) {}
inline void timing_report() {}
inline void timing_reset(const char *, int) {}
inline void timing_active(bool) {}
inline void timing_step(int) {}
inline void timing_comp_proc(const char *, int) {}
inline void timing_depth(int) {}
inline void timing_report_style(const char *) {}
