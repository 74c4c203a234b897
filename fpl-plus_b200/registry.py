"""The plugin dictionaries a PyMIC agent consumes (reference: net/net_dict_seg.py:33-47 key
'UNet2D5_dsbn', loss/loss_dict_seg.py:31-41 keys 'DiceLoss'/'CrossEntropyLoss'); hand them to
``agent.set_net_dict`` / ``agent.set_loss_dict`` (net_run_dsbn/agent_abstract.py:96-110)."""
from .loss import CrossEntropyLoss, DiceLoss
from .net import UNet2D5_dsbn

net_dict = {'UNet2D5_dsbn': UNet2D5_dsbn}
loss_dict = {'DiceLoss': DiceLoss, 'CrossEntropyLoss': CrossEntropyLoss}
