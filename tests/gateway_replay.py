"""Run the UNMODIFIED reference call site ``Gateway.work_flow_fft`` (GT_FFT_v5.py:620-680, from the build-time copy
in oracle/_ref) on .log files, with the hot-path modules bound either to the reference's own files ("ref") or to the
drop-ins of apda-fft_b200/ ("dropin": the sys.path binding of INTEGRATION.md).  Prints fft_dict as JSON.

    python tests/gateway_replay.py ref|dropin <flexible 0|1> <mac> <log file> [<log file> ...]

A separate process per mode, so the two bindings of ``metrics`` / ``utils`` never share an interpreter.
``digidevice`` (the radio driver of the Digi gateway, absent here) is stubbed in sys.modules.
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    mode, flexible, mac, paths = sys.argv[1], bool(int(sys.argv[2])), sys.argv[3], sys.argv[4:]
    ref = os.path.join(ROOT, "oracle", "_ref")
    if mode == "dropin":
        sys.path[:0] = [os.path.join(ROOT, "apda-fft_b200"), ROOT, ref]
    else:
        sys.path[:0] = [ref]
    digi = types.ModuleType("digidevice")
    digi.xbee = types.ModuleType("digidevice.xbee")
    sys.modules["digidevice"] = digi
    sys.modules["digidevice.xbee"] = digi.xbee
    import GT_FFT_v5
    import metrics.fft_iterativa as bound
    gw = object.__new__(GT_FFT_v5.Gateway)
    gw.fft_dict = {}
    gw.is_flexibile_structure = flexible
    gw.append_history = lambda msg: print(msg, file=sys.stderr)
    for path in paths:
        gw.work_flow_fft(mac, path)
    print(json.dumps({"bound": os.path.relpath(bound.__file__, ROOT), "gateway": os.path.relpath(GT_FFT_v5.__file__, ROOT),
                      "fft_dict": gw.fft_dict}))


if __name__ == "__main__":
    main()
