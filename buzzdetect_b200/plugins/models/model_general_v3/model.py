"""B200-native drop-in for the reference's models/model_general_v3/model.py:6-30 (class ModelGeneralV3).

predict(audiosamples) = Dense(13)(embedder.embed(audiosamples)) on raw logits, returned as an object whose
.numpy() is float32 [n_frames, 13] (what src/write/worker.py:69 consumes).  Embedder and head run in ONE engine
call: audio goes to the GPU once and only the [n_frames,13] activations come back.
"""
try:
    import src.config as cfg                       # noqa: F401  (inside a buzzdetect checkout)
    from src.inference.models import BaseModel
except ImportError:
    from buzzdetect_b200 import config as cfg     # noqa: F401
    from buzzdetect_b200.inference.models import BaseModel


import os


class ModelGeneralV3(BaseModel):
    modelname = "model_general_v3"
    # the reference binds the embedder in the class (model.py:8) and has no flag to swap it; BUZZ_B200_EMBEDDER
    # (yamnet | yamnet_k2) is the override hook BASELINE config 5 needs (same head over the Keras-3 embedder)
    embeddername = os.environ.get("BUZZ_B200_EMBEDDER", 'yamnet_k2')
    digits_results = 2

    def initialize(self):
        self.embedder.initialize()
        self.model = self.embedder.model            # the head lives in the same engine as the embedder
        if self.model.n_classes != len(self.config['classes']):
            raise ValueError('head weights and config_model.json disagree on the number of classes')

    def predict(self, audiosamples):
        """Queues the chunk and returns at once; results.numpy() (src/write/worker.py:69) waits for the GPU.  Chunks
        queued while the GPU is busy are computed together in one pass (capi.Engine.submit)."""
        from buzzdetect_b200.results import DeviceResults, as_host_f32
        tk = self.model.submit(as_host_f32(audiosamples), self.embedder.hop_frames)
        return DeviceResults(ticket=tk)

    def predict_pcm(self, pcm, samplerate):
        """The same for a chunk still in its decoded form (int16 / float32, [n] or [n, channels], at the file's rate):
        what the pinned-buffer streamer hands over -- downmix + resample run on the GPU in front of the path."""
        from buzzdetect_b200.results import DeviceResults
        tk = self.model.submit_pcm(pcm, int(samplerate), self.embedder.hop_frames)
        return DeviceResults(ticket=tk)
