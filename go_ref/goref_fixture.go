// Copy to crawler/youtube/goref_fixture.go in the reference checkout: convertVideoToPost (youtube_crawler.go:530-836)
// is unexported, so the YouTube half of the fixture harness has to live inside the package.
package youtube

import (
	"context"
	"encoding/binary"
	"encoding/json"
	"fmt"
	"os"
	"sort"
	"strings"
	"time"

	youtubemodel "github.com/researchaccelerator-hub/telegram-scraper/model/youtube"
)

type fixtureClient struct {
	youtubemodel.YouTubeClient
	chans map[string]*youtubemodel.YouTubeChannel
}

func (f *fixtureClient) GetChannelInfo(ctx context.Context, id string) (*youtubemodel.YouTubeChannel, error) {
	if c, ok := f.chans[id]; ok {
		return c, nil
	}
	return nil, fmt.Errorf("channel %s not cached", id)
}

var thumbKeys = [5]string{"default", "medium", "high", "standard", "maxres"}

// RunFixture: .ytb file (go_ref/dump_batch.py) -> JSONL of convertVideoToPost + json.Marshal, clock injected.
func RunFixture(raw []byte, outPath, linksPath string) {
	le := binary.LittleEndian
	tz := int32(le.Uint32(raw[8:]))
	createdSec, createdNsec := int64(le.Uint64(raw[12:])), int32(le.Uint32(raw[20:]))
	captureSec, captureNsec := int64(le.Uint64(raw[24:])), int32(le.Uint32(raw[32:]))
	labelLen := int(le.Uint32(raw[36:]))
	o := 44
	label := string(raw[o : o+labelLen])
	o += (labelLen + 7) &^ 7
	next := func() []byte {
		n := int(le.Uint64(raw[o:]))
		o += 8
		a := raw[o : o+n]
		o += (n + 7) &^ 7
		return a
	}
	recs, strs, chans, chanStrs := next(), next(), next(), next()
	zone := time.FixedZone("", int(tz))
	time.Local = zone
	fc := &fixtureClient{chans: map[string]*youtubemodel.YouTubeChannel{}}
	var chanIDs []string
	for i := 0; i+64 <= len(chans); i += 64 {
		c := chans[i:]
		so := int(le.Uint32(c[0:]))
		idLen, titleLen, descLen := int(le.Uint16(c[4:])), int(le.Uint16(c[6:])), int(le.Uint32(c[8:]))
		thumbLen, countryLen := int(le.Uint16(c[12:])), int(le.Uint16(c[14:]))
		s := chanStrs[so:]
		id := string(s[:idLen])
		chanIDs = append(chanIDs, id)
		if c[52] == 0 { // not cached: GetChannelInfo fails, convertVideoToPost takes its fallback branch (:808)
			continue
		}
		fc.chans[id] = &youtubemodel.YouTubeChannel{ID: id, Title: string(s[idLen : idLen+titleLen]),
			Description: string(s[idLen+titleLen : idLen+titleLen+descLen]),
			Thumbnails:  map[string]string{"default": string(s[idLen+titleLen+descLen : idLen+titleLen+descLen+thumbLen])},
			Country:     string(s[idLen+titleLen+descLen+thumbLen : idLen+titleLen+descLen+thumbLen+countryLen]),
			SubscriberCount: int64(le.Uint64(c[16:])), ViewCount: int64(le.Uint64(c[24:])), VideoCount: int64(le.Uint64(c[32:])),
			PublishedAt: time.Unix(int64(le.Uint64(c[40:])), int64(int32(le.Uint32(c[48:])))).UTC()}
	}
	cr := &YouTubeCrawler{client: fc, crawlLabel: label}
	created := time.Unix(createdSec, int64(createdNsec)).In(zone)
	capture := time.Unix(captureSec, int64(captureNsec)).In(zone)
	jf, _ := os.Create(outPath)
	lf, _ := os.Create(linksPath)
	defer jf.Close()
	defer lf.Close()
	for i := 0; i+80 <= len(recs); i += 80 {
		r := recs[i:]
		so := int(le.Uint64(r[0:]))
		descLen, chanIdx := int(le.Uint32(r[40:])), int(le.Uint32(r[44:]))
		idLen, titleLen, durLen, langLen := int(le.Uint16(r[48:])), int(le.Uint16(r[50:])), int(le.Uint16(r[52:])), int(le.Uint16(r[54:]))
		s := strs[so:]
		p := 0
		take := func(n int) string { v := string(s[p : p+n]); p += n; return v }
		v := &youtubemodel.YouTubeVideo{ID: take(idLen), Title: take(titleLen), Description: take(descLen), Duration: take(durLen),
			Language: take(langLen), ChannelID: chanIDs[chanIdx], Thumbnails: map[string]string{},
			PublishedAt: time.Unix(int64(le.Uint64(r[8:])), int64(int32(le.Uint32(r[68:])))).UTC(),
			ViewCount: int64(le.Uint64(r[16:])), LikeCount: int64(le.Uint64(r[24:])), CommentCount: int64(le.Uint64(r[32:]))}
		for k := 0; k < 5; k++ {
			if n := int(le.Uint16(r[56+2*k:])); n != 0xFFFF {
				v.Thumbnails[thumbKeys[k]] = take(n)
			}
		}
		post := cr.convertVideoToPost(v)
		post.CreatedAt, post.CaptureTime = created, capture
		line, err := json.Marshal(post)
		if err != nil {
			continue
		}
		jf.Write(append(line, '\n'))
		links := append([]string(nil), post.Outlinks...)
		sort.Strings(links)
		fmt.Fprintf(lf, "%d\t%s\n", i/80, strings.Join(links, ","))
	}
}
