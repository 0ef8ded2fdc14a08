"""GPU: decode once / upload once / slice on the device (DeviceCorpus, aad_extract_indexed) against the
padded-batch path and the reference's per-chunk slicing (ASV_dl_func.py:406-411, 287-293)."""
import wave

import numpy as np
import pytest
import torch

import oracle
from helpers import noise, pad_batch, speech

pytestmark = pytest.mark.gpu
SR = 16000


def _fe(params):
    from audioanalysisdetector_b200 import Frontend
    return Frontend(params, torch.device("cuda:0"))


def _flat(clips, dtype):
    from audioanalysisdetector_b200 import layout_files
    base, total = layout_files([len(c) for c in clips])
    buf = np.zeros(total, dtype=dtype)
    for c, b in zip(clips, base):
        buf[b:b + len(c)] = c
    return buf, base


@pytest.mark.parametrize("kind", ["mfcc", "logmel", "lfcc"])
def test_indexed_chunks_equal_the_padded_batch_bit_for_bit(kind):
    from audioanalysisdetector_b200 import FrontendParams
    dev = torch.device("cuda:0")
    params = {"mfcc": FrontendParams.mfcc(SR, n_mfcc=20, n_delta=2), "logmel": FrontendParams.logmel(SR, n_mels=64),
              "lfcc": FrontendParams.lfcc(SR, n_ceps=13)}[kind]
    fe = _fe(params)
    clips = [speech(1, 5 * SR + 77), noise(2, 3 * SR), speech(3, 70001)]
    buf, base = _flat(clips, np.float32)
    # (file, start, length): plain 2-s chunks, overlapping chunks, a whole file, odd starts (per-sample path)
    table = [(0, 0, 2 * SR), (0, 2 * SR, 2 * SR), (0, SR, 2 * SR), (1, 0, 3 * SR), (2, 0, 70001),
             (2, 12345, 40000), (1, 777, 2 * SR), (0, 4 * SR, SR + 77), (2, 70001 - 9000, 9000)]
    off = np.array([base[f] + s for f, s, n in table], dtype=np.int64)
    ln = np.array([n for f, s, n in table], dtype=np.int32)
    got, nf, st = fe.extract_indexed(torch.from_numpy(buf).to(dev), torch.from_numpy(off), torch.from_numpy(ln))
    wav, lens = pad_batch([clips[f][s:s + n] for f, s, n in table])
    want, nf2, st2 = fe(torch.from_numpy(wav).to(dev), torch.from_numpy(lens).to(dev))
    assert torch.equal(nf, nf2) and torch.equal(st, st2) and int(st.sum()) == 0
    assert got.shape == want.shape and torch.equal(got, want)


def test_indexed_int16_corpus_and_chunk_errors():
    from audioanalysisdetector_b200 import AadError, FrontendParams
    dev = torch.device("cuda:0")
    fe = _fe(FrontendParams.mfcc(SR, n_mfcc=13))
    clips = [np.round(c * 32767).astype(np.int16) for c in (speech(4, 4 * SR), noise(5, 2 * SR + 10))]
    buf, base = _flat(clips, np.int16)
    table = [(0, 0, 2 * SR), (0, 2 * SR, 2 * SR), (1, 10, 2 * SR), (1, 0, 0), (1, 5, 300)]   # empty + short rows
    off = np.array([base[f] + s for f, s, n in table], dtype=np.int64)
    ln = np.array([n for f, s, n in table], dtype=np.int32)
    got, nf, st = fe.extract_indexed(torch.from_numpy(buf).to(dev), torch.from_numpy(off), torch.from_numpy(ln))
    assert st.tolist() == [0, 0, 0, 1, 0]                               # EMPTY row reported, others fine
    for k, (f, s, n) in enumerate(table):
        if n == 0:
            assert float(got[k].abs().max()) == 0.0
            continue
        y = clips[f][s:s + n].astype(np.float32) / 32768.0               # what librosa.load returns
        want = oracle.extract_mfcc_ref(y, SR)
        assert int(nf[k]) == want.shape[1]
        assert np.abs(got[k, :, :want.shape[1]].cpu().numpy() - want).max() <= 1e-3
    with pytest.raises(AadError):                                        # a table that leaves the buffer
        fe.extract_indexed(torch.from_numpy(buf).to(dev), torch.tensor([len(buf) - 100]), torch.tensor([200], dtype=torch.int32))
    with pytest.raises(AadError):
        fe.extract_indexed(torch.from_numpy(buf).to(dev), torch.tensor([-1]), torch.tensor([200], dtype=torch.int32))


def _write_wav(path, y, sr=SR):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.round(y * 32767).astype("<i2").tobytes())


def test_device_corpus_from_files_matches_per_chunk_reference_slicing(tmp_path):
    """prepare_dataframe's 2-s rows (ASV_dl_func.py:287-293) over three files: one decode and one upload
    per file, features equal to the oracle on y[int(cs*sr):min(int(ce*sr), len(y))]."""
    from audioanalysisdetector_b200 import DeviceCorpus, FrontendParams, audio_io, two_second_chunks
    paths = []
    for i, n in enumerate((5 * SR + 900, 2 * SR, SR)):                   # the last one is too short: no rows
        p = tmp_path / f"f{i}.wav"
        _write_wav(p, speech(10 + i, n) if i % 2 else noise(10 + i, n))
        paths.append(str(p))
    corpus = DeviceCorpus("cuda:0")
    rows = []
    for p in paths:
        f = corpus.add(p)
        assert corpus.add(p) == f                                        # decoded once
        rows += [(f, cs, ce) for cs, ce in two_second_chunks(corpus.n_samples[f], corpus.sample_rates[f])]
    assert [r[0] for r in rows] == [0, 0, 1]
    rows.append((0, 4.0, 6.0))                                           # a tail chunk, clipped at len(y)
    off, ln = corpus.table(rows)
    assert ln.tolist() == [2 * SR, 2 * SR, 2 * SR, SR + 900]
    assert corpus.upload().dtype == torch.int16 and corpus.h2d_bytes == corpus.pcm.numel() * 2
    for params, ref in ((FrontendParams.mfcc(SR, n_mfcc=13), oracle.extract_mfcc_ref),
                        (FrontendParams.logmel(SR, n_mels=64), oracle.extract_mel_spectrogram_ref),
                        (FrontendParams.lfcc(SR, n_ceps=13), oracle.extract_lfcc_ref)):
        feats, nf, st = corpus.extract(_fe(params), off, ln)
        assert int(st.sum()) == 0
        for k, (f, cs, ce) in enumerate(rows):
            y, sr = audio_io.load(paths[f])
            want = ref(y, sr, chunk_start=cs, chunk_end=ce)
            got = feats[k].cpu().numpy()
            got = got[:, :nf[k]] if params.layout == 0 else got[:nf[k], :]
            assert got.shape == want.shape and np.abs(got - want).max() <= 1e-3


def test_noise_augmentation_on_device_is_y_plus_factor_times_randn():
    """augment_audio(mode="noise") (ASV_dl_func.py:84-89): y + 1.022 * randn(len(y)), cast to float32;
    only the chosen rows change, and they equal the padded path run on the same noisy clips."""
    from audioanalysisdetector_b200 import DeviceCorpus, FrontendParams
    dev = torch.device("cuda:0")
    corpus = DeviceCorpus(dev)
    clips = [speech(20, 4 * SR), noise(21, 2 * SR + 6)]
    rows = [(corpus.add((clips[0], SR)), 0.0, 2.0), (0, 2.0, 4.0), (corpus.add((clips[1], SR)), None, None)]
    off, ln = corpus.table(rows)
    fe = _fe(FrontendParams.mfcc(SR, n_mfcc=13))
    clean, _, _ = corpus.extract(fe, off, ln)
    clean = clean.clone()
    g = torch.Generator(device=dev).manual_seed(99)
    aug, nf, st = corpus.extract(fe, off, ln, noise_rows=[1, 2], generator=g)
    assert int(st.sum()) == 0 and torch.equal(aug[0], clean[0]) and not torch.equal(aug[1], clean[1])
    g2 = torch.Generator(device=dev).manual_seed(99)
    z = torch.randn(int(ln[1]) + int(ln[2]), dtype=torch.float32, device=dev, generator=g2).cpu().numpy()
    noisy = [clips[0][2 * SR:4 * SR] + np.float32(1.022) * z[:2 * SR], clips[1] + np.float32(1.022) * z[2 * SR:]]
    wav, lens = pad_batch(noisy)
    want, _, _ = fe(torch.from_numpy(wav).to(dev), torch.from_numpy(lens).to(dev))
    for k in range(2):
        T = int(nf[1 + k])
        assert float((aug[1 + k, :, :T] - want[k, :, :T]).abs().max()) <= 1e-3
    # the added noise has the reference's variance: mean power rises by ~ 1.022^2
    y_aug = np.concatenate(noisy)
    assert abs(np.var(y_aug - np.concatenate([clips[0][2 * SR:4 * SR], clips[1]])) / 1.022 ** 2 - 1) < 0.05


def test_long_form_time_split_equals_the_unsplit_extraction():
    """BASELINE config 5 with fewer utterances than GPUs: one utterance split along time over `world`
    ranks (here run one after the other on one GPU, the MAX all-reduce replaced by a hook); the
    concatenated pieces equal the single-pass log-mel (ref=np.max, top_db=80) bit for bit."""
    from audioanalysisdetector_b200 import FrontendParams, long_form_logmel
    dev = torch.device("cuda:0")
    sr, L = 48000, 48000 * 20 + 333
    y = speech(31, L, sr)
    y[: sr] *= 1e-4                                                        # a quiet stretch: the floor bites
    wav = torch.from_numpy(y).to(dev)
    params = FrontendParams.logmel(sr, n_mels=128, n_fft=2048, hop_length=480)
    want, nf, st = _fe(params)(wav[None, :])
    T = int(nf[0])
    assert int(st[0]) == 0 and float(want[0].min()) == -80.0               # top_db active
    for world in (2, 3, 8):
        raws = []
        for r in range(world):                                             # pass 1: local maxima
            raws.append(long_form_logmel(params.replace(ref_type=0, top_db=-1.0), wav, r, world,
                                         local_max_hook=lambda m: m)[0])
        gmax = torch.stack([x.max() for x in raws if x.numel()]).max().reshape(1)
        pieces, covered = [], []
        for r in range(world):
            f, (t0, t1) = long_form_logmel(params, wav, r, world, local_max_hook=lambda m: gmax)
            pieces.append(f)
            covered.append((t0, t1))
        assert covered[0][0] == 0 and covered[-1][1] == T
        got = torch.cat(pieces, dim=1)
        assert got.shape == (128, T) and torch.equal(got, want[0, :, :T])


def test_streamed_upload_of_a_flac_and_wav_corpus(tmp_path):
    """The corpus goes up through two pinned staging buffers (decode of one segment overlaps the copy of the previous
    one); with a tiny segment size files span segments.  The device buffer must equal the back-to-back layout of
    the individually decoded files, whatever the segment size, for FLAC (the ASVspoof container) and WAV alike."""
    import flac_writer as FW
    from audioanalysisdetector_b200 import DeviceCorpus, FrontendParams, audio_io, layout_files
    paths, clips = [], []
    for i, n in enumerate((33000, 4097, 70001, 16000, 5, 52345)):
        pcm = np.round(speech(40 + i, n) * 32767 * 0.9).astype(np.int16) if i % 2 else \
            np.round(noise(40 + i, n) * 32767).astype(np.int16)
        p = tmp_path / (f"LA_{i}.flac" if i % 3 else f"LA_{i}.wav")
        if p.suffix == ".flac":
            p.write_bytes(FW.encode(pcm.astype(np.int64), SR, seed=i))
        else:
            with wave.open(str(p), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(SR)
                w.writeframes(pcm.astype("<i2").tobytes())
        paths.append(str(p))
        clips.append(pcm)
    want, base = _flat(clips, np.int16)
    ups = []
    for stage_bytes, threads in ((64 << 20, None), (16 * 1024, 0), (4096, 3), (8192, 1)):   # decode-ahead pool on / off
        corpus = DeviceCorpus("cuda:0")
        for p in paths:
            corpus.add(p)
        assert corpus.n_samples == [len(c) for c in clips] and corpus._host == [None] * len(paths)   # measured, not decoded
        dev = corpus.upload(stage_bytes=stage_bytes, decode_threads=threads)
        assert dev.dtype == torch.int16 and corpus.base == base
        np.testing.assert_array_equal(dev.cpu().numpy(), want)
        ups.append(corpus)
    # and the features over a chunk table equal the per-file path
    corpus = ups[-1]
    rows = [(0, 0.0, 2.0), (2, 1.0, 3.0), (5, None, None), (3, 0.0, 1.0)]
    off, ln = corpus.table(rows)
    feats, nf, st = corpus.extract(_fe(FrontendParams.mfcc(SR, n_mfcc=13)), off, ln)
    assert int(st.sum()) == 0
    for k, (f, cs, ce) in enumerate(rows):
        y, sr = audio_io.load(paths[f])
        ref = oracle.extract_mfcc_ref(y, sr, chunk_start=cs, chunk_end=ce)
        assert np.abs(feats[k, :, :nf[k]].cpu().numpy() - ref).max() <= 1e-3
