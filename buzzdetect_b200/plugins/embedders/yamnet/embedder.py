"""B200-native drop-in for the reference's embedders/yamnet/embedder.py:14-44 (class EmbedderYamnet, Keras 3).

The reference loads yamnet.keras and sets WaveformFeatures.params.patch_hop_seconds = framehop_s, so any hop is
allowed; the patch step is int(round(100*framehop_s)) STFT frames and the padding hop int32(framehop_s*16000)
(embedders/yamnet/features.py:66-71,99).  Only hops that are whole STFT frames keep time stamps exact
(SURVEY.md section 2b); other hops raise here instead of drifting silently.
"""
import os

try:
    from src.inference.embedding import BaseEmbedder
except ImportError:
    from buzzdetect_b200.inference.embedding import BaseEmbedder


class EmbedderYamnet(BaseEmbedder):
    embeddername = "yamnet"
    framelength_s = 0.96
    digits_time = 2
    samplerate = 16000
    n_embeddings = 1024
    dtype_in = 'float32'

    _mel_variant = "yamnet"

    def initialize(self, engine=None):
        from buzzdetect_b200 import capi
        hop = capi.hop_frames_for(self.framehop_prop)
        if hop < 1 or hop > 96 or abs(hop * 0.01 - self.framehop_s) > 1e-9 or int(self.framehop_s * 16000) != hop * 160:
            raise ValueError(f'framehop_prop={self.framehop_prop} is not a whole number of 10 ms STFT frames')
        self.hop_frames = hop
        if engine is None:
            engine = capi.Engine(device=int(os.environ.get("BUZZ_B200_DEVICE", "0")), embedder=self._mel_variant,
                                 n_slots=int(os.environ.get("BUZZ_B200_SLOTS", "32")))
        self.model = engine

    def embed(self, audio):
        from buzzdetect_b200.results import DeviceResults, as_host_f32
        tk = self.model.submit(as_host_f32(audio), self.hop_frames, want_embeddings=True)
        return DeviceResults(ticket=tk, which="emb")
