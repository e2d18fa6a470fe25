"""cmlpl_b200 -- B200-native (sm_100a) implementation of the CMLPL hot path.

Layout mirrors the reference's module names so its entry points drop in:
    cmlpl_b200.tools.hyper_tools   MirrowCut / ExtractPatches / test_whole / CalAccuracy
    cmlpl_b200.tools.models        BaseNet2 / Normalize / ContrastiveLoss
    cmlpl_b200.hsi_loader          HSIDataSet
    cmlpl_b200.sample_generation   CLI (--dataID --num_label --w --n_PC)
    cmlpl_b200.train               CLI (all train.py flags)
    cmlpl_b200.loss_helper         loss callables
Everything computes through libcmlpl_sm100.so (cmlpl_b200/csrc, C ABI in include/cmlpl.h);
there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
