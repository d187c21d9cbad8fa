"""Parse the option strings of the reference's golden file the way its test does
(transeq/transeq_test.go:64-76: `--frame=6 --table=2 --clean --trim --alternative` style, handed to
go-flags).  Returns keyword arguments for oracle.translate / gotranseq_b200.translate."""


def parse_option_string(s):
    kw = dict(frame="1", table=0, clean=False, alternative=False, trim=False)
    for tok in s.split():
        tok = tok.lstrip("-")
        if "=" in tok:
            k, v = tok.split("=", 1)
        else:
            k, v = tok, None
        if k in ("frame", "f"):
            kw["frame"] = v
        elif k in ("table", "t"):
            kw["table"] = int(v)
        elif k in ("clean", "c"):
            kw["clean"] = True
        elif k in ("alternative", "a"):
            kw["alternative"] = True
        elif k in ("trim", "T"):
            kw["trim"] = True
        elif k in ("numcpu", "n"):
            pass
        else:
            raise ValueError(f"unknown option {tok!r} in {s!r}")
    return kw
