"""init_test_audio - the helper the reference keeps in rawvae/tests.py:13-42 (it is not a test file).

Creates <workdir>/audio_logs, lists the test wavs into <test_audio>.txt, loads and concatenates them (mono, at
sampling_rate), wraps them in a TestDataset and writes test_original.wav. librosa / soundfile are replaced by
audio_io (scipy wav I/O)."""
import os

import numpy as np

from . import audio_io
from .dataset import TestDataset, ToTensor


def init_test_audio(workdir, test_audio, my_test_audio, sampling_rate, segment_length):
    audio_log_dir = workdir / 'audio_logs'
    os.makedirs(audio_log_dir, exist_ok=True)
    test_files = [f for f in my_test_audio.glob('*.wav')]
    with open(audio_log_dir.joinpath(test_audio + '.txt'), 'w') as test_audio_txt:
        test_audio_txt.writelines("{}\n".format(test_file) for test_file in test_files)
    if not test_files:
        raise FileNotFoundError("no .wav files in {}".format(my_test_audio))
    chunks = [audio_io.load_mono(test, sampling_rate)[0] for test in test_files]
    test_dataset_audio = np.concatenate(chunks, axis=0)
    test_dataset = TestDataset(test_dataset_audio, segment_length=segment_length, sampling_rate=sampling_rate,
                               transform=ToTensor())
    audio_io.write_wav(audio_log_dir.joinpath('test_original.wav'), test_dataset_audio, sampling_rate)
    return test_dataset, audio_log_dir
