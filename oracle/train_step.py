"""Oracle: one optimiser step with ``training_all`` semantics (test infrastructure).

Restates PyMIC/pymic/net_run_dsbn/agent_seg.py:459-495: zero_grad; for each
domain d: logits_d = net(x_d, d); L_d = loss(logits_d, y_d[, w_d]); loss =
mean_d L_d (:482 for two domains); backward; Adam(weight_decay) step
(net_run/get_optimizer.py:16-17); MultiStepLR step (:50-54).  The hard-Dice
metric of :472-476 is returned beside the loss.
"""
import torch

from . import losses, unet_dsbn


class OracleTrainer:
    def __init__(self, np_state, net_params, lr=1e-4, weight_decay=1e-5,
                 lr_milestones=None, lr_gamma=0.5, w_dice=1.0, w_ce=0.0):
        self.state = unet_dsbn.to_torch_state(np_state, requires_grad=True)
        self.net_params = net_params
        self.w_dice, self.w_ce = w_dice, w_ce
        # the reference hands *all* parameters to Adam; the ones never used keep
        # grad=None and are skipped by torch.optim.Adam
        self.trainable = [v for v in self.state.values() if v.requires_grad]
        self.opt = torch.optim.Adam(self.trainable, lr, weight_decay=weight_decay)
        self.sched = None
        if lr_milestones is not None:
            self.sched = torch.optim.lr_scheduler.MultiStepLR(self.opt, lr_milestones, lr_gamma)

    def step(self, batches, masks=None):
        """batches: list over domains of (x, soft_y, pixel_weight or None), torch fp32."""
        self.opt.zero_grad()
        total, metrics, logits_all = 0.0, [], []
        for d, (x, y, w) in enumerate(batches):
            logits = unet_dsbn.forward(self.state, x, d, self.net_params, bn_training=True, masks=masks)
            total = total + losses.combined_loss(logits, y, w, self.w_dice, self.w_ce)
            with torch.no_grad():
                metrics.append(losses.hard_dice(logits, y))
            logits_all.append(logits.detach())
        loss = total / len(batches)
        loss.backward()
        self.opt.step()
        if self.sched is not None:
            self.sched.step()
        return loss.item(), metrics, logits_all
