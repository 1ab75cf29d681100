/* ORACLE — test infrastructure.  one translation unit per module (make -C oracle ref compiles this file with
 * -DREFMOD=<name> -DREFMAIN="pipe/modules/<name>/main.c" and -DREF_HAS_* for the callbacks that main.c defines):
 * the reference's module source is included where it lies under /root/reference, its callbacks renamed so that
 * several modules fit into one library, and ref_nodes_<name>() runs them (ref_nodes_driver.h).
 * i-mlv (REFMOD=imlv, -DREF_NO_NODES) is compiled the same way for its init / modify_roi_out: clip header -> image parameters. */
#define REF_CAT_(a, b) a##_##b
#define REF_CAT(a, b)  REF_CAT_(a, b)
#define REF_STR_(a) #a
#define REF_STR(a)  REF_STR_(a)
#define init           REF_CAT(REFMOD, ref_init)
#define cleanup        REF_CAT(REFMOD, ref_cleanup)
#define read_source    REF_CAT(REFMOD, ref_read_source)
#define modify_roi_in  REF_CAT(REFMOD, ref_modify_roi_in)
#define modify_roi_out REF_CAT(REFMOD, ref_modify_roi_out)
#define check_params   REF_CAT(REFMOD, ref_check_params)
#define create_nodes   REF_CAT(REFMOD, ref_create_nodes)
#define commit_params  REF_CAT(REFMOD, ref_commit_params)
#define audio          REF_CAT(REFMOD, ref_audio)
#define write_sink     REF_CAT(REFMOD, ref_write_sink)
#include REFMAIN
#include "ref_nodes_driver.h"

#ifndef REF_HAS_INIT
#define REF_INIT 0
#define REF_CLEANUP 0
#else
#define REF_INIT init
#define REF_CLEANUP cleanup
#endif
#ifndef REF_HAS_ROI_OUT
#define REF_ROI_OUT 0
#else
#define REF_ROI_OUT modify_roi_out
#endif
#ifndef REF_HAS_ROI_IN
#define REF_ROI_IN 0
#else
#define REF_ROI_IN modify_roi_in
#endif

#ifndef REF_NO_NODES /* sources have no create_nodes: only their callbacks are wanted (ref_graph_shim.c binds them) */
int REF_CAT(ref_nodes, REFMOD)(const ref_nodes_in_t *in, char *out, int outsize)
{
  return ref_nodes_run(REF_STR(REFMOD), in, REF_INIT, REF_CLEANUP, REF_ROI_OUT, REF_ROI_IN, create_nodes, out, outsize);
}
#endif
