"""Host-side logic of the Python mirror that needs no GPU (the C-ABI library only has to load)."""
from soundsym_b200 import api


def test_audacity_labels_known_answers(tmp_path):
    """src/sound.rs:561-571 checks tests/vowel.txt for these values; the same three lines in the same format here."""
    p = tmp_path / "labels.txt"
    p.write_text("0.7065779155923718\t0.7619218551399829\to\n0.7619218551399829\t1.0201935730288352\ts\n"
                 "5.4\t5.59353222977394\tning\n3.0\n\nx\ty\n")
    t = api.audacity_labels_to_timestamps(str(p))
    assert len(t) == 6
    assert abs(t[0][0] - 0.7065779155923718) < 1e-10 and abs(t[2][1] - 5.59353222977394) < 1e-10
    assert t[1] == (0.7619218551399829, 1.0201935730288352, "s") and t[2][2] == "ning"
    assert t[3] == (3.0, 0.0, None)    # missing fields: 0.0 / None
    assert t[4] == (0.0, 0.0, None)    # blank line: trim().split('\t') yields one empty field
    assert t[5] == (0.0, 0.0, None)    # unparsable numbers: 0.0


def test_write_wav_rounding(tmp_path):
    import numpy as np
    api._write_wav_i32(str(tmp_path / "a.wav"), np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -2.0, float("nan")]), 44100)
    pcm, sr, bits = api._read_wav_pcm(str(tmp_path / "a.wav"))
    assert bits == 32 and sr == 44100.0
    assert pcm.tolist() == [0, 1073741823, -1073741823, 2147483647, -2147483647, 2147483647, -2147483648, 0]
