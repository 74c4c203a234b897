#!/bin/bash
# Development tool: train-step time with one kernel class removed (FPL_DEBUG_SKIP) = what that class costs in the
# overlapped two-stream schedule.  Numbers are timing-only; results are wrong while a class is skipped.
cd "$(dirname "$0")/.." || exit 1
run() {
  FPL_DEBUG_SKIP="$2" python bench.py --steps 10 --warmup 3 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-28s %.3f ms' % ('$1', d['ms_per_step']))"
}
run baseline ""
run no_bn_bwd "fpl_dsbn_act_bwd_reduce,fpl_dsbn_act_bwd_apply_fin"
run no_bn_fwd "fpl_dsbn_bn_act_fwd"
run no_wgrad "fpl_conv3d_wgrad_tc_tapmajor,fpl_wgrad_tapmajor_to_dw_batch,fpl_conv3d_wgrad_tc_k311,fpl_convt_k2s2_wgrad_tc_tapmajor"
run no_dfold "fpl_conv3d_tc_dfold"
run no_conv_tc "fpl_conv3d_tc"
run no_convt "fpl_convt_k2s2_fwd_tc,fpl_convt_k2s2_dgrad_tc,fpl_convt_k2s2_bwd,fpl_convt_k2s2_wgrad_tc_tapmajor"
run no_head "fpl_head_fwd,fpl_head_dgrad"
run no_stem "fpl_patch9_c8,fpl_conv3d_tc_k311,fpl_conv3d_wgrad_tc_k311"
run no_loss "fpl_dice_ce_reduce_ex,fpl_dice_ce_grad_ex"
run no_adam_fold_scatter "fpl_adam_multi_tensor,fpl_wgrad_tapmajor_to_dw_batch,fpl_grad_scatter_add"
run no_conv_at_all "fpl_conv3d_tc_dfold,fpl_conv3d_tc,fpl_conv3d_wgrad_tc_tapmajor,fpl_convt_k2s2_fwd_tc,fpl_convt_k2s2_dgrad_tc,fpl_convt_k2s2_wgrad_tc_tapmajor,fpl_head_conv_tc,fpl_conv3d_tc_k311,fpl_conv3d_wgrad_tc_k311"
run no_bn_at_all "fpl_dsbn_act_bwd_reduce,fpl_dsbn_act_bwd_apply_fin,fpl_dsbn_bn_act_fwd"
