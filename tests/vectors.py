"""Known-answer vectors for the Decoder hot path.

G* = the reference's own unit tests (file:line cited per vector).
E*/L*/J* = derived behaviour vectors (by code reading).
Expected values are written as the oracle's Rust-style rendering: "Ok(Record {...})" / 'Err("...")'.
"""
TS = "2015-08-05T15:53:45.637824Z"

# --- reference tests -------------------------------------------------------------------------
G1_LINE = r'<23>1 2015-08-05T15:53:45.637824Z testhostname appname 69 42 [origin@123 software="te\st sc\"ript" swVersion="0.0.1"] test message'
G2_LINE = r'<23>1 2015-08-05T15:53:45.637824Z testhostname appname 69 42 [origin@123 software="te\st sc\"ript" swVersion="0.0.1"][master@456 key="value" key2="value2"] test message'
G3_LINE = '{"version":"1.1", "host": "example.org","short_message": "A short message that helps you identify what is going on", "full_message": "Backtrace here\\n\\nmore stuff", "timestamp": 1385053862.3072, "level": 1, "_user_id": 9001, "_some_info": "foo", "_some_env_var": "bar"}'
G9_LINE = "time:1438790025.99\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3"
G10_LINE = "time:[2015-08-05T15:53:45.637824Z]\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3"
G11_LINE = "time:[10/Oct/2000:13:55:36.3 -0700]\tdone:true\tscore:-1\tmean:0.42\tcounter:42\tlevel:3\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3\tmessage:this is a test"
G12_LINE = "time:[5/Aug/2015:15:53:45.637824 -0000]\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3"
G13_LINE = "time:[10/Oct/2000:13:55:36 -0700]\tdone:true\tscore:-1\tmean:0.42\tcounter:42\tlevel:3\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3\tmessage:this is a test"
G14_LINE = "time:[10/Oct/2000:13:55:36 -0700]\tdone_bool:true\tscore_i64:-1\tmean_f64:0.42\tcounter_u64:42\tlevel:3\thost:testhostname\tname1:value1\tname 2: value 2\tn3:v3\tmessage:this is a test"

LTSV_SCHEMA = {"counter": "u64", "score": "i64", "mean": "f64", "done": "bool"}           # ltsv_decoder.rs:371-373
LTSV_SCHEMA_G13 = {"counter": "U64", "score": "I64", "mean": "f64", "done": "bool"}      # :272-277 (mixed case)
LTSV_SUFFIX_G13 = {"u64": "_u64", "i64": "_i64", "F64": "_f64", "Bool": "_bool"}
LTSV_SCHEMA_G14 = {"counter_u64": "U64", "score_i64": "I64", "mean_f64": "f64", "done_bool": "bool"}  # :321-326
LTSV_SUFFIX_G14 = {"u64": "_u64", "i64": "_i64", "f64": "_f64", "bool": "_bool"}

GELF_ERRORS = [  # gelf_decoder.rs:173-205
    ('{"some_key": []}', "Invalid value type in structured data"),
    ('{"timestamp": "a string not a timestamp", "host": "anhostname"}', "Invalid GELF timestamp"),
    ('{some_key = "some_value"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"version":"42"}', "Unsupported GELF version"),
    ('{"level": 8}', "Invalid severity level (too high)"),
]

H = "<13>1 " + TS + " h a p m "  # a valid 6-field header followed by a space

# --- Appendix A.1 (rfc5424_decoder.rs) : (line, error string or None for Ok) -------------------
RFC5424_CASES = [
    ("<34>1 2003-10-11T22:14:15.003Z mymachine.example.com su - ID47 - 'su root' failed", None),   # E1
    ("<165>1 " + TS + ' host app - - [ex@1 iut="3"]', "Missing log message"),                      # E2
    ("<165>1 " + TS + ' host app - - [ex@1 iut="3"] ', None),                                      # E3
    ("<165>1 " + TS + " host app - - [id]", "Missing structured data"),                            # E4
    ("<165>1 " + TS + " host app - - [id] msg", "Missing ] after structured data"),                # E5
    ("<13>1 - host app - - - m", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),  # E6
    ("<13>2 " + TS + " h a p m - x", "Unsupported version"),                                       # E7
    ("<13> " + TS + " h a p m - x", "Unsupported version"),
    ("<13 " + TS + " h a p m - x", "Missing version"),                                             # E8
    ("<>1 " + TS + " h a p m - x", "Invalid priority"),                                            # E9
    ("<256>1 " + TS + " h a p m - x", "Invalid priority"),
    ("<-1>1 " + TS + " h a p m - x", "Invalid priority"),
    ("<1x>1 " + TS + " h a p m - x", "Invalid priority"),
    ("<+7>1 " + TS + " h a p m - x", None),                                                        # E10
    ("<007>1 " + TS + " h a p m - x", None),
    ("abc", "Unsupported BOM"),                                                                    # E11
    ("", "Unsupported BOM"),
    ("\ufeff<13>1 " + TS + " h a p m - x", None),                                                  # E12
    ("\ufeff13>1 " + TS + " h a p m - x", "The priority should be inside brackets"),               # E13
    ("\ufeff", "The priority should be inside brackets"),
    ("<13>1", "Missing timestamp"),                                                                # E14
    ("<13>1  " + TS + " h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),  # E15
    ("<13>1 " + TS, "Missing hostname"),
    ("<13>1 " + TS + " h", "Missing application name"),
    ("<13>1 " + TS + " h a", "Missing process id"),
    ("<13>1 " + TS + " h a p", "Missing message id"),
    ("<13>1 " + TS + " h a p m", "Missing message data"),                                          # E16
    ("<13>1 " + TS + " h a p m ", "Missing log message"),                                          # E17
    (H + "-", None),                                                                               # E18
    (H + "x", "Malformated RFC5424 message"),                                                      # E19
    (H + "[id a=b] m", "Format error in the structured data"),                                     # E20
    (H + '[id a="x]y"] m', None),                                                                  # E21
    (H + '[id a="1"][id2 b="2"]x', "Malformated RFC5424 message"),                                 # E22
    (H + '[id  a="1" ] m', None),                                                                  # E23
    (H + '[id a="1""] m', None),
    (H + '[id a="1"b="2"] m', None),
    (H + '[id a=""] m', None),                                                                     # E24
    (H + '[id ="v"] m', "Format error in the structured data"),                                    # E25
    (H + '[id a="\\\\"] m', None),                                                                 # E26
    (H + "- \u3000hello\u00a0", None),                                                             # E27
    (H + "-abc", None),              # parse_msg(line, 1): anything after '-' is the message
    (H + '[id a="1"]   \t ', None),   # msg None after Unicode trim; full_msg trimmed
    (H + '[ a="1"] m', None),        # empty sd_id
    (H + '[id a="v\\"] m', "Missing ] after structured data"),   # escaped quote never closes
    (H + '[id a="1"]', "Missing log message"),
    (H + '[id a="1" é="2"] m', "Format error in the structured data"),   # non-ASCII cannot start a name
    (H + '[id a="café \\] x"] m', None),
    (H + '[id a\tb="1"] m', "Format error in the structured data"),
    ("<13>1 2015-08-05T15:53:45.637824+25:59 h a p m - x", None),     # UtcOffset range (time >= 0.3.21)
    ("<13>1 2015-08-05T15:53:45.637824+26:00 h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2015-08-05t15:53:45z h a p m - x", None),
    ("<13>1 2015-02-29T15:53:45Z h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2016-02-29T23:59:59.999999999999Z h a p m - x", None),    # >9 fractional digits truncated
    ("<13>1 2016-12-31T23:59:60Z h a p m - x", None),                 # leap second stand-in
    ("<13>1 2016-12-31T22:59:60Z h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2016-12-31T18:59:60-05:00 h a p m - x", None),
    ("<13>1 0000-01-01T00:00:00Z h a p m - x", None),
    ("<13>1 9999-12-31T23:59:59.999999999Z h a p m - x", None),
    ("<13>1 1969-12-31T23:59:59.5Z h a p m - x", None),
    ("<13>1 2015-08-05T15:53:45.Z h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2015-08-05T15:53:45 h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2015-08-05T24:00:00Z h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
    ("<13>1 2015-08-05T15:53:45Zx h a p m - x", "Unable to parse the date from RFC3339 to Unix time in RFC5424 decoder"),
]

# --- Appendix A.2 (ltsv_decoder.rs), no schema ---------------------------------------------------
LTSV_CASES = [
    ("host:h\ttime:1", None),                                                        # L1
    ("time:1\thost:a\thost:b", None),                                                # L2
    ("time:abc\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),  # L3
    ("host:h", "Missing timestamp"),                                                 # L4
    ("time:1", "Missing hostname"),
    ("time:1\thost:h\tlevel:8", "Severity level should be <= 7"),                    # L5
    ("time:1\thost:h\tlevel:x", "Invalid severity level"),
    ("time:1\thost:h\tlevel:+3", None),
    ("time:1\thost:h\tfoo", None),                                                   # L6 (stdout side effect)
    ("time:1\thost:h\tk:v:w", None),                                                 # L7
    ("time:1\thost:h\t:v", None),                                                    # L8
    ("time:inf\thost:h", None),                                                      # L9
    ("time:nan\thost:h", None),
    ("time:-NaN\thost:h", None),
    ("time:1e3\thost:h", None),
    ("time:[1438790025.99]\thost:h", None),                                          # L10
    ("level:9\ttime:x\thost:h", "Severity level should be <= 7"),                    # L11
    ("", "Missing timestamp"),                                                       # L13
    ("foo\tlevel:9\tbar", "Severity level should be <= 7"),   # side effects only for parts before the error
    ("time:1\thost:h\t\t\t", None),                            # empty parts -> three println!
    ("time:[]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:[\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:.5\thost:h", None),
    ("time:5.\thost:h", None),
    ("time:+1.5E+2\thost:h", None),
    ("time:1e\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:0x10\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:1438790025.637824123456789012345\thost:h", None),
    ("time:9007199254740993\thost:h", None),
    ("time:1e400\thost:h", None),
    ("time:1e-400\thost:h", None),
    ("time:2.2250738585072011e-308\thost:h", None),
    ("time:[05/Aug/2015:15:53:45 +0130]\thost:h", None),
    ("time:[0/Aug/2015:15:53:45 +0130]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:[5/aug/2015:15:53:45 +0130]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:[5/Aug/-0044:15:53:45 -0000]\thost:h", None),
    ("time:[5/Aug/2015:15:53:45 0130]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:[5/Aug/2015:15:53:60 +0000]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:[31/Dec/2016:23:59:60 +0000]\thost:h", "Unable to parse the English to Unix timestamp in LTSV decoder"),
    ("time:2016-12-31T23:59:60Z\thost:h", None),
    ("time:[5/Aug/2015:15:53:45.123456789123 +0000]\thost:h", None),
    ("time:1\thost:h\tmessage:a\tmessage:b\té:ü", None),
]

# with LTSV_SCHEMA
LTSV_SCHEMA_CASES = [
    ("time:1\thost:h\tdone:True", "Type error; boolean was expected"),   # L12
    ("time:1\thost:h\tdone:false\tcounter:+7\tscore:+9\tmean:-0.0", None),
    ("time:1\thost:h\tcounter:-1", "Type error; u64 was expected"),
    ("time:1\thost:h\tcounter:18446744073709551615", None),
    ("time:1\thost:h\tcounter:18446744073709551616", "Type error; u64 was expected"),
    ("time:1\thost:h\tscore:-9223372036854775808", None),
    ("time:1\thost:h\tscore:9223372036854775808", "Type error; i64 was expected"),
    ("time:1\thost:h\tscore:", "Type error; i64 was expected"),
    ("time:1\thost:h\tmean:abc", "Type error; f64 was expected"),
    ("time:1\thost:h\tmean:infinity\tmean:1e308\tmean:123456789012345678901234567890", None),
    ("time:1\thost:h\tother:x\tdone:true", None),
]

# --- Appendix A.3 (gelf_decoder.rs) --------------------------------------------------------------
GELF_CASES = [
    ("{}", "Missing hostname"),                                                      # J1
    ("[]", "Empty GELF input"),
    ('"x"', "Empty GELF input"),
    ("", "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h"}', None),                                                          # J2 (ts = now)
    ('{"host":"h","timestamp":1}', None),                                            # J3
    ('{"host":"h","level":-1}', "Invalid severity level"),                           # J4
    ('{"host":"h","level":1.0}', "Invalid severity level"),
    ('{"host":"h","level":"1"}', "Invalid severity level"),
    ('{"host":"h","a":{"b":1}}', "Invalid value type in structured data"),           # J5
    ('{"host":1,"a":[]}', "Invalid value type in structured data"),                  # J6
    ('{"host":"h","x":null,"y":true,"z":-3,"w":2.5,"timestamp":0}', None),           # J7
    ('{"host":"h","_x":1,"x":2,"timestamp":0}', None),                               # J8
    ('{"host":"a","host":"b","timestamp":0}', None),                                 # J9
    ('{"host":"h"} x', "Invalid GELF input, unable to parse as a JSON object"),      # J10
    ('{"host":"h","short_message":"a\nb","timestamp":0}', None),                     # J11
    ('{"version":"1.0","host":"h","timestamp":0}', None),                            # J12
    ('{"version":"2.0","host":"h"}', "Unsupported GELF version"),
    ('{"version":1.1,"host":"h"}', "GELF version must be a string"),
    ('{"host":"h","timestamp":0,"full_message":"\\u00e9\\ud83d\\ude80 \\/ \\b\\f\\r\\t\\"\\\\"}', None),
    ('{"host":"h","timestamp":0,"k":"\\ud83d"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"k":"\\ude80"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"k":"\\x"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"k":"a\tb"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{\n"host":"h","timestamp":0,"k":"a\nb"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"k":"a\\\nb\nc"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"k":"a\nb\\\nc"}', None),   # retry: `\`+LF reads as `\\` then 'n'
    ('{"host":"h","timestamp":0,"a\\u0062c":1,"abc":2}', None),      # duplicate after unescape: last wins
    ('{"host":"h","timestamp":0,"b":1,"a":2,"B":3,"":4,"_":5,"\\u00e9":6,"ab":7}', None),
    ('{"host":"h","timestamp":18446744073709551615}', None),
    ('{"host":"h","timestamp":18446744073709551616}', None),
    ('{"host":"h","timestamp":-9223372036854775808}', None),
    ('{"host":"h","timestamp":-9223372036854775809}', None),
    ('{"host":"h","timestamp":-0}', None),
    ('{"host":"h","timestamp":-0.0}', None),
    ('{"host":"h","timestamp":1e400}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":1e-400}', None),
    ('{"host":"h","timestamp":0e99999999999}', None),
    ('{"host":"h","timestamp":1e99999999999}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":1E-99999999999}', None),
    ('{"host":"h","timestamp":01}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":1.}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":.5}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":-}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":1.7976931348623157e308,"a":123456789012345678901234567890.5e-10}', None),
    ('{"host":"h","timestamp":0.1234567890123456789012345678901234567890}', None),
    ('{"host":"h","timestamp":1385053862.3072,"x":4.9e-324,"y":2.2250738585072014e-308}', None),
    ('{"host":"h","timestamp":0,}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{,"host":"h"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host" "h"}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h" "a":1}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","a":[1,{"b":[true,null,"x\\n"]},-2.5e3],"timestamp":0}', "Invalid value type in structured data"),
    ('{"host":"h","a":[1,]}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","a":tru}', "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","a":nulll}', "Invalid GELF input, unable to parse as a JSON object"),
    (' \t{ "host" : "h" , "timestamp" : 5 } \r', None),
    ('{"host":"h","timestamp":true}', "Invalid GELF timestamp"),
    ('{"host":"h","short_message":1}', "GELF short message must be a string"),
    ('{"host":"h","full_message":null}', "GELF full message must be a string"),
    ('{"host":1}', "GELF host name must be a string"),
    ('{"host":"h","level":7,"timestamp":0}', None),
    ('{"host":"h","level":18446744073709551616}', "Invalid severity level"),
    ('{"hos\\u0074":"esc\\u0061ped host","timestamp":0,"short_message":"tab\\there"}', None),
    ("[" * 127 + "]" * 127, "Empty GELF input"),
    ("[" * 128 + "]" * 128, "Invalid GELF input, unable to parse as a JSON object"),
    ('{"host":"h","timestamp":0,"a":' + "[" * 126 + "]" * 126 + "}", "Invalid value type in structured data"),
    ('{"host":"h","timestamp":0,"a":' + "[" * 127 + "]" * 127 + "}", "Invalid GELF input, unable to parse as a JSON object"),
]

# ---- RFC3164 (decoder/rfc3164_decoder.rs) -------------------------------------------------------------------------
# G16-G26: the reference's own eleven tests.  `partial` = the expected timestamp is built from the CURRENT year
# (ts_from_partial_date_time, utils/test_utils.rs:8-18); the tests below run the decoders with an explicit year instead.
_R3_TAIL = r'''testhostname appname 69 42 [origin@123 software="te\st sc\"ript" swVersion="0.0.1"] test message'''
_R3_MSG = r'''appname 69 42 [origin@123 software="te\st sc\"ript" swVersion="0.0.1"] test message'''
# (name, ref lines, line, expectation) ; expectation None = Err; else dict(fac, sev, date=(y|None, mon, d, h, m, s), host, msg|None, full|None)
RFC3164_GOLDEN = [
    ("G16", "rfc3164_decoder.rs:223-241", "Aug  6 11:15:24 " + _R3_TAIL,
     dict(fac=None, sev=None, date=(None, 8, 6, 11, 15, 24), host="testhostname", msg=_R3_MSG, full=None)),
    ("G17", "rfc3164_decoder.rs:243-262", "<13>Aug  6 11:15:24 " + _R3_TAIL,
     dict(fac=1, sev=5, date=(None, 8, 6, 11, 15, 24), host="testhostname", msg=_R3_MSG, full=None)),
    ("G18", "rfc3164_decoder.rs:264-283", "<13>2020 Aug  6 11:15:24 " + _R3_TAIL,
     dict(fac=1, sev=5, date=(2020, 8, 6, 11, 15, 24), host="testhostname", msg=_R3_MSG, full=None)),
    ("G19", "rfc3164_decoder.rs:285-304", "<13>2020 Aug 6 05:15:24 America/Sao_Paulo " + _R3_TAIL,
     dict(fac=1, sev=5, date=(2020, 8, 6, 8, 15, 24), host="testhostname", msg=_R3_MSG, full=None)),
    ("G20", "rfc3164_decoder.rs:306-325", "Aug  6 11:15:24 UTC " + _R3_TAIL,
     dict(fac=None, sev=None, date=(None, 8, 6, 11, 15, 24), host="testhostname", msg=_R3_MSG, full=None)),
    ("G21", "rfc3164_decoder.rs:327-335", "test message", None),
    ("G22", "rfc3164_decoder.rs:337-345", "Aug  36 11:15:24 " + _R3_TAIL, None),
    ("G23", "rfc3164_decoder.rs:347-369", "testhostname: 2020 Aug  6 11:15:24 UTC: appname 69 42 some test message",
     dict(fac=None, sev=None, date=(2020, 8, 6, 11, 15, 24), host="testhostname", msg="appname 69 42 some test message", full=None)),
    ("G24", "rfc3164_decoder.rs:371-390", "testhostname: 2019 Mar 27 12:09:39: appname: a test message",
     dict(fac=None, sev=None, date=(2019, 3, 27, 12, 9, 39), host="testhostname", msg="appname: a test message", full=None)),
    ("G25", "rfc3164_decoder.rs:392-411", "<13>testhostname: 2019 Mar 27 12:09:39 UTC: appname: test message",
     dict(fac=1, sev=5, date=(2019, 3, 27, 12, 9, 39), host="testhostname", msg="appname: test message", full=None)),
    ("G26", "rfc3164_decoder.rs:413-425", "<13>testhostname: 2019 Mar 27 12:09:39 UTC: appname: test message \n",
     dict(fac=1, sev=5, date=(2019, 3, 27, 12, 9, 39), host="testhostname", msg=None,
          full="<13>testhostname: 2019 Mar 27 12:09:39 UTC: appname: test message")),
]

R3_E_PRI_MALFORMED = "Malformed RFC3164 event: Invalid priority"            # :131
R3_E_PRI = "Invalid priority"                                                # :137
R3_E_CUSTOM = "Malformed RFC3164 event: Invalid timestamp or hostname"       # :120
R3_E_TIME = "Invalid time format"                                            # :158
R3_E_YEAR = "Unable to parse RFC3164 date with year"                         # :178
R3_E_DATE = "Unable to parse the date in RFC3164 decoder"                    # :211
R3_E_PANIC = "(the reference panics here: index out of bounds, rfc3164_decoder.rs:64)"

# Derived behaviour vectors (read off rfc3164_decoder.rs; each: line, expected error or None for Ok).  Decoded with year 2026.
RFC3164_CASES = [
    ("", R3_E_CUSTOM),
    ("<", R3_E_PRI_MALFORMED),
    ("<13", R3_E_PRI_MALFORMED),
    ("<>Aug 6 11:15:24 h m", R3_E_PRI),
    ("<<13>Aug 6 11:15:24 h m", None),                 # trim_start_matches('<') drops every '<'
    ("<+13>Aug 6 11:15:24 h m", None),                 # u8::from_str accepts a '+'
    ("<013>Aug 6 11:15:24 h m", None),
    ("<256>Aug 6 11:15:24 h m", R3_E_PRI),
    ("<-1>Aug 6 11:15:24 h m", R3_E_PRI),
    ("<1 3>Aug 6 11:15:24 h m", R3_E_PRI),
    ("<13>", R3_E_CUSTOM),
    (" <13>Aug 6 11:15:24 h m", R3_E_CUSTOM),          # no '<' at byte 0: "<13>Aug" is the month token
    ("Aug 6 11:15:24 h", None),                        # 4 tokens, message = ""
    ("Aug 6 11:15:24", R3_E_CUSTOM),                   # 3 tokens only
    ("Aug 6 11:15:24 UTC", R3_E_PANIC),                # the zone eats the 4th token: `_log_tokens[0]` on an empty Vec
    ("2020 Aug 6 11:15:24 UTC", R3_E_PANIC),
    ("2020 Aug 6 11:15:24", R3_E_PANIC),               # the with-year form consumes all four tokens
    ("2020 Aug 6 11:15:24 h", None),
    ("2020 Aug 6 11:15:24 h m", None),
    ("Aug 6 11:15:24 UTC h", None),
    ("Aug 6 11:15:24 utc h m", None),                  # get_by_name is exact: "utc" is the hostname
    ("Aug 6 11:15:24 Europe/Paris h m", None),
    ("Aug 6 11:15:24 Europe/paris h m", None),
    ("Aug 06 11:15:24 h m", None),                     # [day padding:none] still takes two digits
    ("Aug 0 11:15:24 h m", R3_E_CUSTOM),
    ("Aug 006 11:15:24 h m", R3_E_CUSTOM),
    ("Aug 31 11:15:24 h m", None),
    ("Sep 31 11:15:24 h m", R3_E_CUSTOM),
    ("Feb 29 11:15:24 h m", R3_E_CUSTOM),              # 2026 is not a leap year
    ("2024 Feb 29 11:15:24 h m", None),
    ("aug 6 11:15:24 h m", R3_E_CUSTOM),               # month names are case-sensitive
    ("AUG 6 11:15:24 h m", R3_E_CUSTOM),
    ("August 6 11:15:24 h m", R3_E_CUSTOM),
    ("Aug 6 24:00:00 h m", R3_E_CUSTOM),
    ("Aug 6 23:59:60 h m", R3_E_CUSTOM),
    ("Aug 6 23:59:59 h m", None),
    ("Aug 6 1:15:24 h m", R3_E_CUSTOM),                # [hour] is exactly two digits
    ("Aug 6 11:15:24.5 h m", R3_E_CUSTOM),
    ("Aug 6 11:15 h m", R3_E_CUSTOM),
    ("+2020 Aug 6 11:15:24 h m", None),                # [year] sign:automatic
    ("-2020 Aug 6 11:15:24 h m", None),
    ("-0000 Feb 29 11:15:24 h m", None),
    ("02020 Aug 6 11:15:24 h m", R3_E_CUSTOM),
    ("202 Aug 6 11:15:24 h m", R3_E_CUSTOM),
    ("9999 Dec 31 23:59:59 h m", None),
    ("0000 Jan 1 00:00:00 Asia/Tokyo h m", None),
    ("Aug\t6 11:15:24 h　m  n\r\no ", None),      # split_whitespace is Unicode White_Space
    ("Aug 6 11:15:24 h m\u200bn", None),               # U+200B is not White_Space
    ("Aug 6 11:15:24 h été m", None),
    ("Aug 6 11:15:24 h a  b", None),                   # the message is re-joined with single spaces
    ("Aug 6 11:15:24 h a b   ", None),
    ("   Aug 6 11:15:24 h a b", None),
    ("h: 2020 Aug 6 11:15:24: m", None),
    ("h: Aug 6 11:15:24: m", None),                    # the custom form without a year
    ("h: 2020 Aug 6 11:15:24 UTC junk: m", None),      # tokens after the date / zone are ignored
    ("h: 2020 Aug 6 11:15:24 junk: m", None),
    ("h: Aug 6: m", R3_E_TIME),                        # two date tokens
    ("h: Aug 6 x: m", R3_E_YEAR),                      # three: the without-year parse fails, the with-year one needs four
    ("h: Aug 6 x y: m", R3_E_DATE),
    ("h: Aug 6 11:15:24", R3_E_CUSTOM),                # only two ": " pieces
    ("h: a: m", R3_E_TIME),
    ("h: : m", R3_E_TIME),
    (": 2020 Aug 6 11:15:24: ", None),                 # empty hostname, empty message
    ("a b: 2020 Aug 6 11:15:24: m: n: o ", None),      # the hostname may hold spaces; the message keeps ": " and is not trimmed
    ("h:  2020 Aug 6 11:15:24: m", None),
    ("h: 2020 Aug 6 11:15:24:m: n", R3_E_DATE),        # "11:15:24:m" is the time token
    ("<13>h: 2020 Aug 6 11:15:24 America/New_York: m", None),
    ("2021 Mar 14 02:30:00 America/New_York h m", None),   # a local time the zone skips (parity unpinned: offset before the jump)
    ("2021 Nov 7 01:30:00 America/New_York h m", None),    # a local time that occurs twice (parity unpinned: the first)
    ("2050 Jul 1 12:00:00 Europe/Paris h m", None),        # past the explicit transitions: POSIX footer rule
    ("1890 Jul 1 12:00:00 Europe/Paris h m", None),        # local mean time
    ("2020 Aug 6 11:15:24 Etc/GMT+5 h m", None),
    ("2020 Aug 6 11:15:24 EST5EDT h m", None),
    ("2020 Aug 6 11:15:24 posixrules h m", None),          # not an IANA name: stays the hostname
    ("Aug 6 11:15:24 America/Argentina/ComodRivadavia h m", None),
]
