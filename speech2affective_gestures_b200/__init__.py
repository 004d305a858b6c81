"""s2ag-b200: B200-native implementation of the Speech2AffectiveGestures GAN-step hot path.

Public surface mirrors the reference for this path:
  speech2affective_gestures_b200.net.multimodal_context_net_v2  -- PoseGenerator, PoseGeneratorTriModal,
        AffDiscriminator, ConvDiscriminatorTriModal, WavEncoder, MFCCEncoder, TextEncoderTCN, AffEncoder
  speech2affective_gestures_b200.net.tcn                         -- TemporalBlock, TemporalConvNet
  speech2affective_gestures_b200.processor_v2                    -- Processor (train / generate entry points)
All compute runs in libs2ag_b200.so (hand-written sm_100a CUDA behind the C ABI of include/s2ag.h).
"""
__version__ = "0.1.0"
