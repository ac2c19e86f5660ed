"""B200-native batched likelihood hot path of The Payne.

Public surface mirrors the reference for this path:
    thepayne_b200.fitting.likelihood.likelihood      (lnlikefn / lnlike / lnlike_batch)
    thepayne_b200.fitting.genmod.GenMod              (genspec / genphot / genphot_scaled)
    thepayne_b200.predict.predictspec.ANN, PayneSpecPredict
    thepayne_b200.predict.predictsed.FastPayneSEDPredict
backed by the C ABI of include/payne_b200.h (thepayne_b200/libpayne_b200.so).
"""
__version__ = '0.1.0'
