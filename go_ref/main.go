// Command goref feeds a packed batch (go_ref/dump_batch.py) through the REAL reference code and writes what it
// produced: telegramhelper.ParseMessage + json.Marshal for .tgb files, YouTubeCrawler.convertVideoToPost for .ytb
// files.  See README.md.  Lives in the reference module as ./cmd/goref; cannot be built in the engine's own image
// (no Go toolchain there).
package main

import (
	"encoding/binary"
	"encoding/json"
	"flag"
	"fmt"
	"os"
	"sort"
	"strings"
	"time"

	"github.com/researchaccelerator-hub/telegram-scraper/common"
	"github.com/researchaccelerator-hub/telegram-scraper/crawler"
	ytcrawler "github.com/researchaccelerator-hub/telegram-scraper/crawler/youtube"
	"github.com/researchaccelerator-hub/telegram-scraper/model"
	"github.com/researchaccelerator-hub/telegram-scraper/state"
	"github.com/researchaccelerator-hub/telegram-scraper/telegramhelper"
	"github.com/zelenin/go-tdlib/client"
)

// ---- the packed batch (include/tgingest.h), little endian ----------------------------------------------------
type tgRec struct {
	ID, ChatID, MediaAlbumID int64
	StrOff                   uint64
	Date, ViewCount, ShareCount int32
	ChanIdx, TextLen, AltLen    uint32
	MediaLen, HandleLen         uint16
	ContentType, Flags          uint8
	Reserved                    uint16
}
type entity struct {
	Offset, Length int32
	URLOff         uint32
	URLLen         uint16
	Type, Reserved uint8
}
type reaction struct {
	EmojiOff           uint32
	EmojiLen, Reserved uint16
	Count              int32
}
type comment struct {
	TextOff, TextLen, HandleOff uint32
	HandleLen                   uint16
	Flags, Reserved             uint8
	ViewCount, ReplyCount       int32
	ReactStart, ReactCount      uint32
}
type tgChan struct {
	StrOff                      uint32
	TitleLen, NameLen, UserLen  uint16
	Reserved                    uint16
	Reserved2                   uint32
	MemberCount, PostCount, ViewCount int64
}
type fileCfg struct {
	Version                 uint32
	TZ                      int32
	CreatedSec              int64
	CreatedNsec             int32
	CaptureSec              int64
	CaptureNsec             int32
	LabelLen                uint32
}

type reader struct {
	b []byte
	o int
}

func (r *reader) array() []byte {
	n := int(binary.LittleEndian.Uint64(r.b[r.o:]))
	r.o += 8
	a := r.b[r.o : r.o+n]
	r.o += (n + 7) &^ 7
	return a
}
func decode[T any](raw []byte, size int) []T {
	out := make([]T, len(raw)/size)
	for i := range out {
		if err := binary.Read(strings.NewReader(string(raw[i*size:(i+1)*size])), binary.LittleEndian, &out[i]); err != nil {
			panic(err)
		}
	}
	return out
}
func u32s(raw []byte) []uint32 {
	out := make([]uint32, len(raw)/4)
	for i := range out {
		out[i] = binary.LittleEndian.Uint32(raw[4*i:])
	}
	return out
}

// content types, in the order of the TGI_CT_* enum
const (
	ctNone = iota
	ctText
	ctVideo
	ctPhoto
	ctAnimation
	ctAnimatedEmoji
	ctPoll
	ctGiveaway
	ctPaidMedia
	ctSticker
	ctGiveawayWinners
	ctGiveawayCompleted
	ctVideoNote
	ctDocument
	ctAudio
	ctVoiceNote
	ctOther
)

// ---- stubs of the two interfaces ParseMessage talks to ----------------------------------------------------------
type stubTD struct {
	crawler.TDLibClient // every method the path does not call stays nil (and would panic loudly)
	share               map[[2]int64]int32
	poster              map[int64]string // sender chat id -> title
	threads             map[[2]int64][]*client.Message
}

func (s *stubTD) GetMessage(req *client.GetMessageRequest) (*client.Message, error) {
	return &client.Message{Id: req.MessageId, ChatId: req.ChatId,
		InteractionInfo: &client.MessageInteractionInfo{ForwardCount: s.share[[2]int64{req.ChatId, req.MessageId}]}}, nil
}
func (s *stubTD) GetChat(req *client.GetChatRequest) (*client.Chat, error) {
	return &client.Chat{Id: req.ChatId, Title: s.poster[req.ChatId]}, nil
}
func (s *stubTD) GetMessageThreadHistory(req *client.GetMessageThreadHistoryRequest) (*client.Messages, error) {
	if req.FromMessageId != 0 { // second page: nothing left
		return &client.Messages{}, nil
	}
	m := s.threads[[2]int64{req.ChatId, req.MessageId}]
	return &client.Messages{TotalCount: int32(len(m)), Messages: m}, nil
}

type stubSM struct {
	state.StateManagementInterface
	last *model.Post
}

func (s *stubSM) StorePost(channelID string, post model.Post) error { s.last = &post; return nil }

func str(b []byte, off uint64, n uint32) string { return string(b[off : off+uint64(n)]) }

func formatted(text string, ents []entity, aux []byte) *client.FormattedText {
	ft := &client.FormattedText{Text: text}
	for _, e := range ents {
		te := &client.TextEntity{Offset: e.Offset, Length: e.Length}
		switch e.Type {
		case 1:
			te.Type = &client.TextEntityTypeTextUrl{Url: string(aux[e.URLOff : e.URLOff+uint32(e.URLLen)])}
		case 2:
			te.Type = &client.TextEntityTypeMention{}
		case 3:
			te.Type = &client.TextEntityTypeUrl{}
		default:
			te.Type = &client.TextEntityTypeBold{}
		}
		ft.Entities = append(ft.Entities, te)
	}
	return ft
}

func remote(id string) *client.File { return &client.File{Remote: &client.RemoteFile{Id: id}} }

func reactionsOf(rs []reaction, aux []byte) *client.MessageReactions {
	out := &client.MessageReactions{}
	for _, r := range rs {
		out.Reactions = append(out.Reactions, &client.MessageReaction{
			Type: &client.ReactionTypeEmoji{Emoji: string(aux[r.EmojiOff : r.EmojiOff+uint32(r.EmojiLen)])}, TotalCount: r.Count})
	}
	return out
}

func runTelegram(raw []byte, outPath, linksPath string) {
	rd := &reader{b: raw, o: 4}
	var cfgf fileCfg
	_ = binary.Read(strings.NewReader(string(raw[4:4+40])), binary.LittleEndian, &cfgf)
	rd.o = 4 + 40
	label := string(raw[rd.o : rd.o+int(cfgf.LabelLen)])
	rd.o += (int(cfgf.LabelLen) + 7) &^ 7
	recs := decode[tgRec](rd.array(), 64)
	strs := rd.array()
	entOff, ents := u32s(rd.array()), decode[entity](rd.array(), 16)
	reactOff, reacts := u32s(rd.array()), decode[reaction](rd.array(), 12)
	commentOff, comments := u32s(rd.array()), decode[comment](rd.array(), 32)
	aux := rd.array()
	chans, chanStrs := decode[tgChan](rd.array(), 40), rd.array()

	zone := time.FixedZone("", int(cfgf.TZ))
	time.Local = zone // ParseMessage formats time.Unix(..) in the process-local zone
	created := time.Unix(cfgf.CreatedSec, 0).UTC().Truncate(time.Second)
	capture := time.Unix(cfgf.CaptureSec, int64(cfgf.CaptureNsec)).In(zone)
	cfg := common.CrawlerConfig{SkipMediaDownload: true, MaxComments: 1 << 20}

	jf, _ := os.Create(outPath)
	lf, _ := os.Create(linksPath)
	defer jf.Close()
	defer lf.Close()
	for i, r := range recs {
		so := r.StrOff
		text := str(strs, so, r.TextLen)
		alt := str(strs, so+uint64(r.TextLen), r.AltLen)
		media := str(strs, so+uint64(r.TextLen)+uint64(r.AltLen), uint32(r.MediaLen))
		handle := str(strs, so+uint64(r.TextLen)+uint64(r.AltLen)+uint64(r.MediaLen), uint32(r.HandleLen))
		var ft *client.FormattedText
		if r.Flags&1 != 0 {
			ft = formatted(text, ents[entOff[i]:entOff[i+1]], aux)
		}
		senderID := int64(i) + 7_000_000_000
		td := &stubTD{share: map[[2]int64]int32{{r.ChatID, r.ID}: r.ShareCount}, poster: map[int64]string{senderID: handle},
			threads: map[[2]int64][]*client.Message{}}
		msg := &client.Message{Id: r.ID, ChatId: r.ChatID, Date: r.Date, MediaAlbumId: client.JsonInt64(r.MediaAlbumID),
			SenderId: &client.MessageSenderChat{ChatId: senderID}}
		cs := comments[commentOff[i]:commentOff[i+1]]
		msg.InteractionInfo = &client.MessageInteractionInfo{ViewCount: r.ViewCount, ForwardCount: r.ShareCount,
			Reactions: reactionsOf(reacts[reactOff[i]:reactOff[i+1]], aux)}
		if len(cs) > 0 {
			msg.InteractionInfo.ReplyInfo = &client.MessageReplyInfo{ReplyCount: int32(len(cs))}
			var thread []*client.Message
			for k, c := range cs {
				cm := &client.Message{Id: int64(k+1) << 20, ChatId: r.ChatID,
					Content:  &client.MessageText{Text: &client.FormattedText{Text: string(aux[c.TextOff : c.TextOff+c.TextLen])}},
					SenderId: &client.MessageSenderChat{ChatId: senderID + int64(k+1)*1_000_000_000},
					InteractionInfo: &client.MessageInteractionInfo{ViewCount: c.ViewCount,
						ReplyInfo: &client.MessageReplyInfo{ReplyCount: c.ReplyCount}}}
				if c.Flags&1 != 0 {
					cm.InteractionInfo.Reactions = reactionsOf(reacts[c.ReactStart:c.ReactStart+c.ReactCount], aux)
				}
				td.poster[senderID+int64(k+1)*1_000_000_000] = string(aux[c.HandleOff : c.HandleOff+uint32(c.HandleLen)])
				thread = append(thread, cm)
			}
			td.threads[[2]int64{r.ChatID, r.ID}] = thread
		}
		switch r.ContentType {
		case ctNone:
		case ctText:
			msg.Content = &client.MessageText{Text: ft}
		case ctVideo:
			v := &client.MessageVideo{Caption: ft}
			if media != "" { // TGI_RF_PANIC records (thumbnail without caption / file) are built by hand in the tests, not here
				v.Video = &client.Video{Video: remote(media), Thumbnail: &client.Thumbnail{File: remote("")}}
			}
			msg.Content = v
		case ctPhoto:
			msg.Content = &client.MessagePhoto{Caption: ft}
		case ctAnimation:
			msg.Content = &client.MessageAnimation{Caption: ft}
		case ctAnimatedEmoji:
			msg.Content = &client.MessageAnimatedEmoji{Emoji: alt}
		case ctPoll:
			msg.Content = &client.MessagePoll{Poll: &client.Poll{Question: &client.FormattedText{Text: alt}}}
		case ctGiveaway:
			if alt == "giveawayPrizePremium" {
				msg.Content = &client.MessageGiveaway{Prize: &client.GiveawayPrizePremium{}}
			} else {
				msg.Content = &client.MessageGiveaway{Prize: &client.GiveawayPrizeStars{}}
			}
		case ctPaidMedia:
			msg.Content = &client.MessagePaidMedia{Caption: &client.FormattedText{Text: alt}}
		case ctSticker:
			msg.Content = &client.MessageSticker{}
		case ctGiveawayWinners:
			msg.Content = &client.MessageGiveawayWinners{}
		case ctGiveawayCompleted:
			msg.Content = &client.MessageGiveawayCompleted{}
		case ctVideoNote:
			msg.Content = &client.MessageVideoNote{VideoNote: &client.VideoNote{Video: remote(media)}}
		case ctDocument:
			msg.Content = &client.MessageDocument{Caption: ft, Document: &client.Document{FileName: alt, Document: remote(media)}}
		case ctAudio:
			msg.Content = &client.MessageAudio{Caption: ft}
		case ctVoiceNote:
			msg.Content = &client.MessageVoiceNote{Caption: ft}
		default: // ctOther: the generator uses these four
			switch alt {
			case "messageLocation":
				msg.Content = &client.MessageLocation{}
			case "messageContact":
				msg.Content = &client.MessageContact{}
			case "messageDice":
				msg.Content = &client.MessageDice{}
			default:
				msg.Content = &client.MessageStory{}
			}
		}
		ch := chans[r.ChanIdx]
		cso := uint64(ch.StrOff)
		title := str(chanStrs, cso, uint32(ch.TitleLen))
		name := str(chanStrs, cso+uint64(ch.TitleLen), uint32(ch.NameLen))
		user := str(chanStrs, cso+uint64(ch.TitleLen)+uint64(ch.NameLen), uint32(ch.UserLen))
		sg := &client.Supergroup{}
		if user != "" {
			sg.Usernames = &client.Usernames{ActiveUsernames: []string{user}}
		}
		sm := &stubSM{}
		post, err := telegramhelper.ParseMessage("fixture", msg, &client.Chat{Id: r.ChatID, Title: title}, sg,
			&client.SupergroupFullInfo{MemberCount: int32(ch.MemberCount)}, int(ch.PostCount), int(ch.ViewCount), name, td, sm, cfg)
		if err != nil || sm.last == nil { // failed / skipped: no line (crawl/runner.go:1199-1214, tdutils.go:419-421)
			continue
		}
		_ = post
		p := *sm.last
		p.CreatedAt, p.CaptureTime, p.CrawlLabel = created, capture, label // the clock and the sink's label are injected
		line, merr := json.Marshal(p)                                      // state/storageproviders.go:276-282
		if merr != nil {
			continue
		}
		jf.Write(append(line, '\n'))
		links := append([]string(nil), p.Outlinks...)
		sort.Strings(links)
		fmt.Fprintf(lf, "%d\t%s\n", i, strings.Join(links, ","))
	}
}

func main() {
	in := flag.String("in", "", "packed batch (.tgb / .ytb)")
	out := flag.String("out", "", "JSONL output")
	links := flag.String("links", "", "per-record outlinks output")
	flag.Parse()
	raw, err := os.ReadFile(*in)
	if err != nil {
		panic(err)
	}
	switch string(raw[:4]) {
	case "TGB1":
		runTelegram(raw, *out, *links)
	case "YTB1":
		ytcrawler.RunFixture(raw, *out, *links) // crawler/youtube/goref_fixture.go (convertVideoToPost is unexported)
	default:
		panic("unknown file kind")
	}
}
