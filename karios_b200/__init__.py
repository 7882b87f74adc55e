"""karios_b200 -- B200-native KLT matching hot path of KARIOS.

Public surface (mirrors karios.matcher / karios.core of the reference):
    karios_b200.matcher.klt.klt_tracker, .KLT
    karios_b200.matcher.zncc_service.ZNCCService
    karios_b200.core.configuration.KLTConfiguration
    karios_b200.core.image.ArrayRaster, .DeviceRaster
The arithmetic lives in karios_b200/_lib/libkarios_b200.so (include/karios_b200.h).
"""
__version__ = "0.1.0"
