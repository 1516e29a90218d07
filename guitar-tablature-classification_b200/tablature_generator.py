"""Drop-in for the FEATURE side of the reference's ``tablature_generator.py`` (class TablatureImageGenerator, :474-666):
``segment_audio`` and the numeric part of ``audio_to_cqt_image`` on a B200.  The CNN, the matplotlib rendering, the MP3
decode (pydub) and the PIL tablature drawing are out of scope.
"""
from __future__ import annotations

import os

import numpy as np

from gtc_b200.inference import TabCnnFrontEnd


class TablatureImageGenerator:
    """Feature front-end only: same method names and arguments as the reference where they exist."""

    def __init__(self, model_path=None, device=None):
        self.model_path = model_path
        self._fe = TabCnnFrontEnd(device=device)
        os.makedirs("temp_spectrograms", exist_ok=True)                      # :498

    def segment_audio(self, audio_file, segment_duration=3.0, sr=22050, overlap=0.5):
        """tablature_generator.py:637-666."""
        if int(sr) != self._fe.sr:
            self._fe = TabCnnFrontEnd(sr=int(sr))
        return self._fe.segment_audio(audio_file, segment_duration, sr, overlap)

    def audio_to_cqt_image(self, audio_file, output_path=None, sr=22050, hop_length=512):
        """tablature_generator.py:599-635 without the drawing: saves the (84, T) dB array specshow would receive as
        ``<output_path>.npy`` and returns that path."""
        if output_path is None:
            output_path = os.path.join("temp_spectrograms", f"{os.path.basename(audio_file)}_spectrogram.png")
        if (int(sr), int(hop_length)) != (self._fe.sr, self._fe.hop_length):
            self._fe = TabCnnFrontEnd(sr=int(sr), hop_length=int(hop_length))
        C_mag = self._fe.audio_to_cqt_db(audio_file)
        np.save(output_path + ".npy", np.asfortranarray(C_mag))
        return output_path + ".npy"

    def segment_features(self, audio_file, segment_duration=3.0, overlap=0.5):
        """All segments of generate_tablature_from_mp3's loop (:871-884) in one batch: ([n_seg, 84, T] device tensor, times)."""
        y = self._fe.load(audio_file)
        return self._fe.cqt_db_segments(y, segment_duration, overlap)
