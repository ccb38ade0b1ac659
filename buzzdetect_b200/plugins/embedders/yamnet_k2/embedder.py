"""B200-native drop-in for the reference's embedders/yamnet_k2/embedder.py:5-37 (class YamnetK2).

Same class attributes, same constructor, same initialize()/embed() contract; the TFSMLayer over the Keras-2
SavedModel is replaced by buzzdetect_b200's CUDA engine.  Differences, all deliberate:
  * framehop_prop 0.5 works (the reference raises KeyError: its halfhop graph's output key is
    'global_average_pooling2d_1');
  * embed() returns an object with .numpy() instead of a tf.Tensor.
"""
import os

try:                                    # inside a buzzdetect checkout
    from src.inference.embedding import BaseEmbedder
except ImportError:                     # inside buzzdetect_b200
    from buzzdetect_b200.inference.embedding import BaseEmbedder


class YamnetK2(BaseEmbedder):
    embeddername = "yamnet"            # sic: the reference's attribute value (embedder.py:7)
    framelength_s = 0.96
    digits_time = 2
    samplerate = 16000
    n_embeddings = 1024
    dtype_in = 'float32'

    _mel_variant = "yamnet_k2"

    def initialize(self, engine=None):
        if self.framehop_prop not in (1, 0.5):
            raise ValueError('For Keras 2 YAMNet, framehop_prop must be 1 or 0.5')
        from buzzdetect_b200 import capi
        self.hop_frames = capi.hop_frames_for(self.framehop_prop)
        if engine is None:
            engine = capi.Engine(device=int(os.environ.get("BUZZ_B200_DEVICE", "0")), embedder=self._mel_variant,
                                 n_slots=int(os.environ.get("BUZZ_B200_SLOTS", "32")))
        self.model = engine

    def embed(self, audiosamples):
        from buzzdetect_b200.results import DeviceResults, as_host_f32
        tk = self.model.submit(as_host_f32(audiosamples), self.hop_frames, want_embeddings=True)
        return DeviceResults(ticket=tk, which="emb")
